// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT.  C ABI over the oracle port
// so tests/ and bench.py's cpu_baseline leg can drive it through ctypes.  The
// stage-level entry points mirror oracle/ref_tap.cc one to one (op_* vs ref_*) so
// the same test code can compare port, reference and CUDA results.
#include "oracle_port.hpp"

#include <atomic>
#include <climits>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>
#include <chrono>

using namespace oport;

namespace {
struct op_index {
  sr_index idx;
  std::vector<int> unitigs_lengths;
  std::vector<std::string> unitigs_sequences;
};
struct op_aligner {
  op_index* oi;
  params    p;
  std::vector<int64_t> groups; std::vector<int32_t> offsets; std::vector<uint32_t> lis;
  std::vector<int64_t> cint; std::vector<double> cdbl; std::vector<int64_t> info_off;
  std::vector<int32_t> kinfo, binfo;
};

struct read_rec { std::string name, seq; };
// FASTA / FASTQ reader with the behaviour of Jellyfish's whole_sequence_parser as the
// reference uses it (create_mega_reads.cc:52-58): header without '>'/'@', line breaks removed.
struct read_stream {
  std::vector<std::string> paths; size_t next_path = 0; std::unique_ptr<std::ifstream> cur;
  bool next(read_rec& r) {
    while(true) {
      if(!cur) {
        if(next_path >= paths.size()) return false;
        cur.reset(new std::ifstream(paths[next_path++]));
        if(!cur->good()) throw std::runtime_error("Can't open read file");
      }
      int c = cur->peek();
      while(c == '\n' || c == '\r') { cur->get(); c = cur->peek(); }
      if(c == EOF) { cur.reset(); continue; }
      std::string header, line;
      if(c == '>') {
        cur->get(); std::getline(*cur, header); r.seq.clear();
        for(c = cur->peek(); c != '>' && c != EOF; c = cur->peek()) { std::getline(*cur, line); r.seq += line; }
      } else if(c == '@') {
        cur->get(); std::getline(*cur, header); r.seq.clear();
        for(c = cur->peek(); c != '+' && c != EOF; c = cur->peek()) { std::getline(*cur, line); r.seq += line; }
        if(c == '+') {
          std::getline(*cur, line);
          size_t q = 0;
          while(q < r.seq.size() && cur->good()) { std::getline(*cur, line); q += line.size(); }
        }
      } else throw std::runtime_error("Unsupported format");
      r.name = header.substr(0, header.find_first_of(" \t\n\v\f\r"));
      return true;
    }
  }
};
}

extern "C" {

void* op_index_create(const char* sr_fasta, unsigned min_size, unsigned max_size) {
  try {
    std::unique_ptr<op_index> r(new op_index);
    r->idx.append_fasta(sr_fasta);
    r->idx.build(min_size, max_size);
    return r.release();
  } catch(std::exception& e) { std::cerr << "op_index_create: " << e.what() << std::endl; return 0; }
}
void op_index_destroy(void* p) { delete (op_index*)p; }
uint64_t op_index_n(void* p) { return ((op_index*)p)->idx.n(); }
uint64_t op_index_nseq(void* p) { return ((op_index*)p)->idx.srs.size(); }
uint64_t op_index_sa_size(void* p) { return ((op_index*)p)->idx.sa.size(); }
void op_index_sa(void* p, uint64_t* out) { auto& v = ((op_index*)p)->idx.sa; memcpy(out, v.data(), v.size() * 8); }
void op_index_counts(void* p, uint64_t* out) { auto& v = ((op_index*)p)->idx.counts; memcpy(out, v.data(), v.size() * 8); }
void op_index_seq_starts(void* p, uint64_t* out) { auto& v = ((op_index*)p)->idx.starts; memcpy(out, v.data(), v.size() * 8); }
// text as one code (0..3) per byte
void op_index_text_codes(void* p, uint8_t* out) { auto& v = ((op_index*)p)->idx.text; memcpy(out, v.data(), v.size()); }
void op_index_search_k(void* p, const uint64_t* mers, uint64_t q, unsigned kk, uint64_t* index_out, uint64_t* nb_out) {
  op_index* oi = (op_index*)p;
  for(uint64_t i = 0; i < q; ++i) oi->idx.search_k(mers[i], kk, index_out[i], nb_out[i]);
}
void op_index_search(void* p, const uint64_t* mers, uint64_t q, uint64_t* index_out, uint64_t* nb_out) {
  const sr_index& idx = ((op_index*)p)->idx;
  for(uint64_t i = 0; i < q; ++i) idx.search(mers[i], index_out[i], nb_out[i]);
}
uint32_t op_lis(const int32_t* pairs, uint32_t n, double a, double b, double C, uint32_t window, uint32_t* out) {
  if(window < 1) return UINT32_MAX;
  std::vector<std::pair<int,int>> X(n);
  for(uint32_t i = 0; i < n; ++i) X[i] = std::make_pair(pairs[2 * i], pairs[2 * i + 1]);
  const auto res = chain(X, a, b, C, window);
  for(size_t i = 0; i < res.size(); ++i) out[i] = res[i];
  return res.size();
}
int op_index_set_unitigs_lengths(void* p, const int32_t* lens, uint64_t n) {
  ((op_index*)p)->unitigs_lengths.assign(lens, lens + n);
  return 0;
}
void* op_aligner_create(void* p, double stretch_factor, double stretch_constant, double stretch_cap,
                        uint32_t window_size, int forward, int max_match, int max_count,
                        double matching_mers, double matching_bases, uint32_t unitigs_k) {
  op_aligner* a = new op_aligner;
  a->oi = (op_index*)p;
  a->p.stretch_factor = stretch_factor; a->p.stretch_constant = stretch_constant; a->p.stretch_cap = stretch_cap;
  a->p.window_size = window_size; a->p.forward = forward; a->p.max_match = max_match; a->p.max_count = max_count;
  a->p.matching_mers = matching_mers; a->p.matching_bases = matching_bases; a->p.unitigs_k = unitigs_k;
  return a;
}
void op_aligner_destroy(void* p) { delete (op_aligner*)p; }

static void flatten(op_aligner* a, const std::vector<mer_lists>& groups, const std::vector<coords>& cs) {
  a->groups.clear(); a->offsets.clear(); a->lis.clear();
  for(const auto& ml : groups) {
    const int64_t g[5] = { ml.sr, (int64_t)ml.fwd.offsets.size(), (int64_t)ml.bwd.offsets.size(),
                           (int64_t)ml.fwd.lis.size(), (int64_t)ml.bwd.lis.size() };
    a->groups.insert(a->groups.end(), g, g + 5);
    for(const auto& x : ml.fwd.offsets) { a->offsets.push_back(x.first); a->offsets.push_back(x.second); }
    for(const auto& x : ml.bwd.offsets) { a->offsets.push_back(x.first); a->offsets.push_back(x.second); }
    a->lis.insert(a->lis.end(), ml.fwd.lis.begin(), ml.fwd.lis.end());
    a->lis.insert(a->lis.end(), ml.bwd.lis.begin(), ml.bwd.lis.end());
  }
  a->cint.clear(); a->cdbl.clear(); a->info_off.assign(1, 0); a->kinfo.clear(); a->binfo.clear();
  for(const auto& c : cs) {
    const int64_t v[14] = { c.rs, c.re, c.qs, c.qe, c.nb_mers, c.pb_cons, c.sr_cons, c.pb_cover, c.sr_cover,
                            (int64_t)c.rl, (int64_t)c.ql, c.rn, c.sr, c.use_bwd_name };
    a->cint.insert(a->cint.end(), v, v + 14);
    a->cdbl.push_back(c.stretch); a->cdbl.push_back(c.offset); a->cdbl.push_back(c.avg_err);
    a->kinfo.insert(a->kinfo.end(), c.kmers_info.begin(), c.kmers_info.end());
    a->binfo.insert(a->binfo.end(), c.bases_info.begin(), c.bases_info.end());
    a->info_off.push_back(a->kinfo.size());
  }
}

// one read; with max_match the groups' lists are the post-discard state, as in the reference
int op_align_read(void* p, const char* seq, uint64_t len) {
  op_aligner* a = (op_aligner*)p;
  try {
    std::vector<mer_lists> groups; std::vector<coords> cs;
    align_read(a->oi->idx, std::string(seq, len), a->p, a->p.unitigs_k ? &a->oi->unitigs_lengths : nullptr, groups, cs);
    flatten(a, groups, cs);
    return 0;
  } catch(std::exception& e) { std::cerr << "op_align_read: " << e.what() << std::endl; return -1; }
}
uint64_t op_res_ngroups(void* p)  { return ((op_aligner*)p)->groups.size() / 5; }
uint64_t op_res_noffsets(void* p) { return ((op_aligner*)p)->offsets.size() / 2; }
uint64_t op_res_nlis(void* p)     { return ((op_aligner*)p)->lis.size(); }
uint64_t op_res_ncoords(void* p)  { return ((op_aligner*)p)->cint.size() / 14; }
uint64_t op_res_ninfo(void* p)    { return ((op_aligner*)p)->kinfo.size(); }
void op_res_copy(void* p, int64_t* groups, int32_t* offsets, uint32_t* lis, int64_t* cint, double* cdbl,
                 int64_t* info_off, int32_t* kinfo, int32_t* binfo) {
  op_aligner* a = (op_aligner*)p;
  if(groups)   memcpy(groups, a->groups.data(), a->groups.size() * 8);
  if(offsets)  memcpy(offsets, a->offsets.data(), a->offsets.size() * 4);
  if(lis)      memcpy(lis, a->lis.data(), a->lis.size() * 4);
  if(cint)     memcpy(cint, a->cint.data(), a->cint.size() * 8);
  if(cdbl)     memcpy(cdbl, a->cdbl.data(), a->cdbl.size() * 8);
  if(info_off) memcpy(info_off, a->info_off.data(), a->info_off.size() * 8);
  if(kinfo)    memcpy(kinfo, a->kinfo.data(), a->kinfo.size() * 4);
  if(binfo)    memcpy(binfo, a->binfo.data(), a->binfo.size() * 4);
}

// kmers_info known-answer hook (reference tests/test_kmers_info.cc): feeds positions one by one
// and writes the vectors' state after each add_mer as "len v0 v1 ... len b0 b1 ..." rows.
int op_kmers_info_trace(const char* name, const int32_t* ul, uint32_t n_ul, uint32_t unitigs_k, uint32_t k,
                        const int32_t* positions, uint32_t npos, int32_t* out, uint32_t out_cap) {
  const std::vector<uint32_t> u = parse_sr_name(name);
  const std::vector<int> lens(ul, ul + n_ul);
  std::vector<int> mers, bases;
  kmers_info_state st(mers, bases, u, unitigs_k, k, &lens);
  uint32_t w = 0;
  auto dump = [&]() {
    if(w + 2 + 2 * mers.size() > out_cap) return false;
    out[w++] = mers.size(); for(int v : mers) out[w++] = v;
    out[w++] = bases.size(); for(int v : bases) out[w++] = v;
    return true;
  };
  if(!dump()) return -1;
  for(uint32_t i = 0; i < npos; ++i) { st.add_mer(positions[i]); if(!dump()) return -1; }
  return w;
}

int op_sr_overlap(const char* a, const char* b) { return sr_overlap(parse_sr_name(a), parse_sr_name(b)); }

// Whole-path drivers: mode 0 = create_mega_reads (create_mega_reads.cc:95-167), 1 = jf_aligner coords
// (jf_aligner.cc:161-233, compact format, no header).  `unitigs_is_fasta` selects -u vs -l.
// Returns the number of read bases processed (< 0 on error); *align_seconds gets the time of the
// per-read phase only (the reference's "create mega reads" timer).
// --fine-mer for the next op_run calls (0 = no fine pass); kept out of op_run's long argument list
static unsigned g_fine_mer = 0;
void op_set_fine_mer(unsigned k) { g_fine_mer = k; }

int64_t op_run(int mode, const char* sr_fasta, const char* reads_path, const char* unitigs_path, int unitigs_is_fasta,
               const char* out_path, unsigned mer, unsigned psa_min, unsigned threads, uint64_t max_reads,
               double stretch_factor, double stretch_constant, double stretch_cap, int forward, int max_match,
               int max_count, double mers_matching_pct, double bases_matching_pct, unsigned unitigs_k,
               double overlap_play, double errors, double density, double min_length, int bases, int tiling, int trim,
               double* index_seconds, double* align_seconds) {
  try {
    op_index oi;
    const auto t0 = std::chrono::steady_clock::now();
    oi.idx.append_fasta(sr_fasta);
    std::cerr << "compute_psa " << oi.idx.srs.size() << ' ' << oi.idx.n() << '\n';
    // create_mega_reads.cc:131-132: the suffix array keeps suffixes down to min(fine mer, psa-min) bases
    oi.idx.build(g_fine_mer ? std::min(g_fine_mer, psa_min) : psa_min, mer);
    const auto t1 = std::chrono::steady_clock::now();
    if(unitigs_path && *unitigs_path) {
      std::ifstream is(unitigs_path);
      if(!is.good()) throw std::runtime_error("Failed to open unitigs file");
      if(unitigs_is_fasta) {                                   // misc.cc:31-37
        std::string hdr, seq;
        while(std::getline(is, hdr)) {
          std::getline(is, seq);
          oi.unitigs_sequences.push_back(seq);
          oi.unitigs_lengths.push_back(seq.size());
        }
      } else {                                                 // misc.cc:11-19
        std::string u; unsigned len;
        while(is >> u >> len) oi.unitigs_lengths.push_back(len);
      }
    }
    params p;
    p.stretch_factor = stretch_factor; p.stretch_constant = stretch_constant; p.stretch_cap = stretch_cap;
    p.forward = forward; p.max_match = max_match; p.max_count = max_count;
    p.matching_mers = mers_matching_pct / 100.0; p.matching_bases = bases_matching_pct / 100.0;
    p.unitigs_k = oi.unitigs_lengths.empty() ? 0 : unitigs_k;
    p.overlap_play = overlap_play; p.errors = errors; p.density = density; p.min_length = min_length;
    p.bases = bases; p.tiling = tiling; p.trim = trim;

    read_stream rs; rs.paths.push_back(reads_path);
    std::ofstream os(out_path);
    if(!os.good()) throw std::runtime_error("Failed to open output file");
    std::mutex in_mutex, out_mutex;
    std::atomic<int64_t> total_bases(0);
    uint64_t nread = 0;
    bool failed = false;
    const auto t2 = std::chrono::steady_clock::now();
    auto worker = [&]() {
      std::vector<read_rec> job(100);
      std::string out;
      std::vector<mer_lists> groups; std::vector<coords> cs, fine;
      while(true) {
        size_t filled = 0;
        {
          std::lock_guard<std::mutex> lock(in_mutex);
          while(filled < job.size() && (max_reads == 0 || nread < max_reads) && rs.next(job[filled])) { ++filled; ++nread; }
        }
        if(!filled) break;
        for(size_t i = 0; i < filled; ++i) {
          try {
            align_read(oi.idx, job[i].seq, p, p.unitigs_k ? &oi.unitigs_lengths : nullptr, groups, cs);
            if(g_fine_mer) {                                    // create_mega_reads.cc:64-68, jf_aligner.cc:144-148
              fine_align_read(oi.idx, job[i].seq, p, g_fine_mer, p.unitigs_k ? &oi.unitigs_lengths : nullptr, cs, fine);
              cs.swap(fine);
            }
            if(mode == 0)
              mega_reads_for_read(oi.idx, cs, job[i].name, job[i].seq.size(), p, oi.unitigs_lengths,
                                  oi.unitigs_sequences.empty() ? nullptr : &oi.unitigs_sequences, out);
            else
              print_coords(oi.idx, cs, job[i].name, job[i].seq.size(), out);
          } catch(std::exception& e) { std::cerr << "op_run: " << e.what() << std::endl; failed = true; }
          total_bases += job[i].seq.size();
        }
        if(out.size() > (1 << 20)) { std::lock_guard<std::mutex> lock(out_mutex); os << out; out.clear(); }
      }
      std::lock_guard<std::mutex> lock(out_mutex);
      os << out;
    };
    std::vector<std::thread> th;
    for(unsigned t = 0; t < std::max(1u, threads); ++t) th.emplace_back(worker);
    for(auto& t : th) t.join();
    const auto t3 = std::chrono::steady_clock::now();
    if(index_seconds) *index_seconds = std::chrono::duration<double>(t1 - t0).count();
    if(align_seconds) *align_seconds = std::chrono::duration<double>(t3 - t2).count();
    return failed ? -1 : (int64_t)total_bases;
  } catch(std::exception& e) { std::cerr << "op_run: " << e.what() << std::endl; return -1; }
}

} // extern "C"
