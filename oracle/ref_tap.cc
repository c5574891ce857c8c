// TEST INFRASTRUCTURE ONLY.  Stage-level taps into the REAL reference code
// (compiled from /root/reference where it lies; see oracle/Makefile).  Exposes a
// C ABI so tests can pull the reference's intermediate results through ctypes:
//   - suffix array + prefix counts   (src_psa/mer_sa_imp.hpp:197-267)
//   - k-mer search                   (src_psa/mer_sa_imp.hpp:369-479)
//   - chaining                       (src_lis/lis_align.hpp:139-204)
//   - per-read hit lists, chains and coords
//                                    (src_jf_aligner/coarse_aligner.cc:42-141,
//                                     src_jf_aligner/pb_aligner.cc:11-82)
// Built into oracle/_ref/libref_tap.so with -fno-access-control so that private
// members (PSA::m_mer_counts) can be read without touching reference sources.
#include <src_jf_aligner/superread_parser.hpp>
#include <src_jf_aligner/coarse_aligner.hpp>
#include <src_lis/lis_align.hpp>
#include <src_jf_aligner/jf_aligner.hpp>
#include <src_jf_aligner/fine_aligner.hpp>
#include <sstream>

using align_pb::coarse_aligner;

struct ref_index {
  sequence_psa psa;
  unsigned int min_size, max_size;
  std::vector<int> unitigs_lengths;
};

struct ref_aligner {
  ref_index*                       idx;
  std::unique_ptr<coarse_aligner>  aligner;
  std::unique_ptr<coarse_aligner::thread> th;
  // last result, flattened
  std::vector<int64_t> groups;   // per group: sr_index, n_fwd, n_bwd, lis_fwd, lis_bwd
  std::vector<int32_t> offsets;  // (pb, sr) pairs, fwd then bwd for each group in order
  std::vector<uint32_t> lis;     // fwd lis then bwd lis for each group in order
  std::vector<int64_t> cint;     // per coords 14 ints
  std::vector<double>  cdbl;     // per coords 3 doubles
  std::vector<int64_t> info_off; // CSR offsets into kinfo/binfo (ncoords + 1)
  std::vector<int32_t> kinfo, binfo;
};

extern "C" {

void* ref_index_create(const char* sr_fasta, unsigned int min_size, unsigned int max_size, unsigned int threads) {
  try {
    ref_index* r = new ref_index;
    jellyfish::mer_dna::k(max_size);
    r->psa.append_fasta(sr_fasta);
    r->min_size = min_size;
    r->max_size = max_size;
    r->psa.compute_psa(min_size, max_size, threads);
    return r;
  } catch(std::exception& e) {
    std::cerr << "ref_index_create: " << e.what() << std::endl;
    return 0;
  }
}
void ref_index_destroy(void* p) { delete (ref_index*)p; }
uint64_t ref_index_n(void* p) { return ((ref_index*)p)->psa.sequence_size(); }
uint64_t ref_index_nseq(void* p) { return ((ref_index*)p)->psa.nb_sequences(); }
uint64_t ref_index_sa_size(void* p) { ref_index* r = (ref_index*)p; return r->psa.sequence_size() - r->min_size + 1; }
void ref_index_sa(void* p, uint64_t* out) {
  ref_index* r = (ref_index*)p;
  const uint64_t n = ref_index_sa_size(p);
  for(uint64_t i = 0; i < n; ++i) out[i] = (*r->psa.m_sa)[i];
}
// 4^min_size + 1 entries
void ref_index_counts(void* p, uint64_t* out) {
  ref_index* r = (ref_index*)p;
  const size_t nb = ((size_t)1 << (2 * r->min_size)) + 1;
  memcpy(out, r->psa.m_sa->m_mer_counts.get(), nb * sizeof(uint64_t));
}
// 2-bit text words as the reference packed them
uint64_t ref_index_text_words(void* p) { return ((ref_index*)p)->psa.m_sequence.size(); }
void ref_index_text(void* p, uint64_t* out) {
  ref_index* r = (ref_index*)p;
  memcpy(out, r->psa.m_sequence.data(), r->psa.m_sequence.size() * sizeof(uint64_t));
}
void ref_index_seq_starts(void* p, uint64_t* out) { // nseq + 1
  ref_index* r = (ref_index*)p;
  for(size_t i = 0; i < r->psa.m_offsets.size(); ++i) out[i] = r->psa.m_offsets[i].sequence;
}
// search k-mers given as integers (first base most significant); k = max_size
void ref_index_search(void* p, const uint64_t* mers, uint64_t q, uint64_t* index_out, uint64_t* nb_out) {
  ref_index* r = (ref_index*)p;
  const unsigned int k = r->max_size;
  jellyfish::mer_dna::k(k);
  for(uint64_t i = 0; i < q; ++i) {
    jellyfish::mer_dna m;
    for(unsigned int j = 0; j < k; ++j)
      m.shift_left((int)((mers[i] >> (2 * (k - 1 - j))) & 3));
    auto res = r->psa.m_sa->search(mer_dna_ptr<jellyfish::mer_dna>(m), k);
    nb_out[i]    = res.first;
    index_out[i] = res.first ? res.second : 0;
  }
}

// the same search with a pattern of kk bases, as the fine pass issues it (fine_aligner.cc:13-15 ->
// superread_parser.hpp:183-192): mer type short_mer_type, Psize = kk
void ref_index_search_k(void* p, const uint64_t* mers, uint64_t q, unsigned int kk, uint64_t* index_out, uint64_t* nb_out) {
  ref_index* r = (ref_index*)p;
  short_mer_type::k(kk);
  for(uint64_t i = 0; i < q; ++i) {
    short_mer_type m;
    for(unsigned int j = 0; j < kk; ++j)
      m.shift_left((int)((mers[i] >> (2 * (kk - 1 - j))) & 3));
    auto res = r->psa.m_sa->search(mer_dna_ptr<short_mer_type>(m), kk);
    nb_out[i]    = res.first;
    index_out[i] = res.first ? res.second : 0;
  }
}

// chaining on (pb, sr) int pairs; returns chain length, indices into out (size >= n)
uint32_t ref_lis(const int32_t* pairs, uint32_t n, double a, double b, double C, uint32_t window, uint32_t* out) {
  std::vector<std::pair<int,int> > X(n);
  for(uint32_t i = 0; i < n; ++i) X[i] = std::make_pair(pairs[2 * i], pairs[2 * i + 1]);
  align_pb::lis_buffer_type L;
  std::vector<unsigned int> res;
  lis_align::affine_capped accept_mer(a, b, C);
  lis_align::linear        accept_seq(a);
  const unsigned int len = lis_align::indices(X.cbegin(), X.cend(), L, res, window, accept_mer, accept_seq);
  for(uint32_t i = 0; i < len; ++i) out[i] = res[i];
  return len;
}

int ref_index_set_unitigs_lengths(void* p, const int32_t* lens, uint64_t n) {
  ref_index* r = (ref_index*)p;
  r->unitigs_lengths.assign(lens, lens + n);
  return 0;
}

void* ref_aligner_create(void* p, double stretch_factor, double stretch_constant, double stretch_cap,
                         uint32_t window_size, int forward, int max_match, int max_count,
                         double matching_mers, double matching_bases, uint32_t unitigs_k) {
  ref_index* r = (ref_index*)p;
  ref_aligner* a = new ref_aligner;
  a->idx = r;
  jellyfish::mer_dna::k(r->max_size);
  a->aligner.reset(new coarse_aligner(r->psa, r->max_size, stretch_factor, stretch_constant, stretch_cap,
                                      window_size, forward, max_match,
                                      max_count ? max_count : std::numeric_limits<int>::max(),
                                      matching_mers, matching_bases));
  if(unitigs_k) a->aligner->unitigs_lengths(&r->unitigs_lengths, unitigs_k);
  a->th.reset(new coarse_aligner::thread(*a->aligner));
  return a;
}
void ref_aligner_destroy(void* p) { delete (ref_aligner*)p; }

// diagnostic: run the fine pass on the read last given to ref_align_read and print, for every window
// of super-read `sr`, its bounds and first hits (fine_aligner.hpp:49-58, fine_aligner.cc:7-36)
void ref_fine_dump(void* p, const char* seq, uint64_t len, unsigned int fine_k, int64_t sr) {
  ref_aligner* a = (ref_aligner*)p;
  const std::string s(seq, len);
  short_mer_type::k(fine_k);
  align_pb::fine_aligner fa(a->idx->psa, fine_k, a->aligner->unitigs_lengths_, a->aligner->unitigs_k_);
  align_pb::fine_aligner::thread th(fa);
  short_parse_sequence parser(s);
  th.align_sequence(parser, s.size(), a->th->coords());
  const frag_lists::frag_info* base = a->idx->psa.m_headers.data();
  for(const auto& it : th.frags_pos_)
    for(const auto& w : it.second) {
      if(w.ml.frag - base != sr) continue;
      std::cerr << "ref window sr " << sr << " begin " << w.begin << " end " << w.end << " fwd " << w.ml.fwd.offsets.size() << " bwd " << w.ml.bwd.offsets.size() << " first bwd:";
      for(size_t i = 0; i < w.ml.bwd.offsets.size() && i < 6; ++i) std::cerr << ' ' << w.ml.bwd.offsets[i].first << ':' << w.ml.bwd.offsets[i].second;
      std::cerr << " lis bwd " << w.ml.bwd.lis.size() << std::endl;
    }
}

// align one read; groups are emitted sorted by super-read index
int ref_align_read(void* p, const char* seq, uint64_t len) {
  ref_aligner* a = (ref_aligner*)p;
  const std::string s(seq, len);
  parse_sequence parser(s);
  a->th->align_sequence_max(parser, s.size());
  const auto& fp = a->th->frags_pos();
  const frag_lists::frag_info* base = a->idx->psa.m_headers.data();

  std::vector<std::pair<int64_t, const align_pb::mer_lists*> > order;
  for(const auto& it : fp) order.push_back(std::make_pair((int64_t)(it.second.frag - base), &it.second));
  std::sort(order.begin(), order.end());
  a->groups.clear(); a->offsets.clear(); a->lis.clear();
  for(const auto& o : order) {
    const auto& ml = *o.second;
    a->groups.push_back(o.first);
    a->groups.push_back(ml.fwd.offsets.size());
    a->groups.push_back(ml.bwd.offsets.size());
    a->groups.push_back(ml.fwd.lis.size());
    a->groups.push_back(ml.bwd.lis.size());
    for(const auto& x : ml.fwd.offsets) { a->offsets.push_back(x.first); a->offsets.push_back(x.second); }
    for(const auto& x : ml.bwd.offsets) { a->offsets.push_back(x.first); a->offsets.push_back(x.second); }
    for(auto x : ml.fwd.lis) a->lis.push_back(x);
    for(auto x : ml.bwd.lis) a->lis.push_back(x);
  }

  const auto& coords = a->th->coords();
  a->cint.clear(); a->cdbl.clear(); a->info_off.assign(1, 0); a->kinfo.clear(); a->binfo.clear();
  for(const auto& c : coords) {
    const int64_t v[14] = { c.rs, c.re, c.qs, c.qe, c.nb_mers, c.pb_cons, c.sr_cons, c.pb_cover, c.sr_cover,
                            (int64_t)c.rl, (int64_t)c.ql, c.rn, (int64_t)(c.qfrag - base),
                            c.name_u == &c.qfrag->bwd };
    a->cint.insert(a->cint.end(), v, v + 14);
    a->cdbl.push_back(c.stretch); a->cdbl.push_back(c.offset); a->cdbl.push_back(c.avg_err);
    a->kinfo.insert(a->kinfo.end(), c.kmers_info.begin(), c.kmers_info.end());
    a->binfo.insert(a->binfo.end(), c.bases_info.begin(), c.bases_info.end());
    a->info_off.push_back(a->kinfo.size());
  }
  return 0;
}
uint64_t ref_res_ngroups(void* p)  { return ((ref_aligner*)p)->groups.size() / 5; }
uint64_t ref_res_noffsets(void* p) { return ((ref_aligner*)p)->offsets.size() / 2; }
uint64_t ref_res_nlis(void* p)     { return ((ref_aligner*)p)->lis.size(); }
uint64_t ref_res_ncoords(void* p)  { return ((ref_aligner*)p)->cint.size() / 14; }
uint64_t ref_res_ninfo(void* p)    { return ((ref_aligner*)p)->kinfo.size(); }
void ref_res_copy(void* p, int64_t* groups, int32_t* offsets, uint32_t* lis, int64_t* cint, double* cdbl,
                  int64_t* info_off, int32_t* kinfo, int32_t* binfo) {
  ref_aligner* a = (ref_aligner*)p;
  if(groups)   memcpy(groups, a->groups.data(), a->groups.size() * sizeof(int64_t));
  if(offsets)  memcpy(offsets, a->offsets.data(), a->offsets.size() * sizeof(int32_t));
  if(lis)      memcpy(lis, a->lis.data(), a->lis.size() * sizeof(uint32_t));
  if(cint)     memcpy(cint, a->cint.data(), a->cint.size() * sizeof(int64_t));
  if(cdbl)     memcpy(cdbl, a->cdbl.data(), a->cdbl.size() * sizeof(double));
  if(info_off) memcpy(info_off, a->info_off.data(), a->info_off.size() * sizeof(int64_t));
  if(kinfo)    memcpy(kinfo, a->kinfo.data(), a->kinfo.size() * sizeof(int32_t));
  if(binfo)    memcpy(binfo, a->binfo.data(), a->binfo.size() * sizeof(int32_t));
}

} // extern "C"
