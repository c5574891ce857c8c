// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT.
// CPU restatement ("port") of the reference's create_mega_reads / jf_aligner hot
// path, written from the reference's behaviour (file:line cited at each
// function) as plain sequential C++.  It is the checker for the CUDA path in
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the product
// library (pacbio_b200/csrc) never includes, links or calls it.
//
// Parity status: PINNED.  tests/test_oracle_*.py check this port against
//  (1) the reference's own golden vectors (tests/golden/*, taken from
//      /root/reference/tests/{test_kmers_info.cc,test_lis_align.cc,
//      aligner_output/*}), and
//  (2) the reference itself compiled from /root/reference (oracle/_ref, see
//      oracle/Makefile) on seeded synthetic inputs, stage by stage.
#ifndef ORACLE_PORT_HPP
#define ORACLE_PORT_HPP
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace oport {

// ---- index ---------------------------------------------------------------
struct super_read {
  std::string           name;        // forward name (header line after '>')
  std::vector<uint32_t> fwd_u;       // (id << 1) | ori(R=1), empty if the name does not parse
  std::vector<uint32_t> bwd_u;       // reversed path
  std::string           bwd_name;
  uint64_t              start;       // offset in concatenated text
  uint32_t              len;
};

struct sr_index {
  unsigned              m = 13, k = 17;        // psa-min, mer
  std::vector<uint8_t>  text;                   // codes 0..3, concatenated, no separators
  std::vector<super_read> srs;
  std::vector<uint64_t> starts;                 // nseq + 1
  std::vector<uint64_t> sa;                     // n - m + 1 positions
  std::vector<uint64_t> counts;                 // 4^m + 1
  uint64_t n() const { return text.size(); }

  void append_fasta(const std::string& path);  // superread_parser.cc:12-46
  void build(unsigned m_, unsigned k_);        // mer_sa_imp.hpp:197-253,352-366
  uint64_t kmer_at(uint64_t pos, unsigned len) const;   // padded with A past the end
  void search(uint64_t mer, uint64_t& index_out, uint64_t& nb_out) const; // mer_sa_imp.hpp:369-479
  // the same search with a pattern of kk <= k bases (the fine pass looks up shorter mers in the same SA)
  void search_k(uint64_t mer, unsigned kk, uint64_t& index_out, uint64_t& nb_out) const;
  bool locate_k(uint64_t x, unsigned kk, uint32_t& sr, int32_t& off) const;
  // SA entry -> (super-read, 1-based offset); false if the k-mer straddles two sequences
  bool locate(uint64_t x, uint32_t& sr, int32_t& off) const;  // superread_parser.hpp:110-134
};

// ---- per-read alignment ----------------------------------------------------
struct params {
  double   stretch_factor = 1.3, stretch_constant = 10, stretch_cap = 10000;
  uint32_t window_size = 1;
  bool     forward = true, max_match = false;
  int      max_count = 5000;                   // 0 => INT_MAX
  double   matching_mers = 0.0, matching_bases = 0.17;   // already divided by 100
  uint32_t unitigs_k = 0;                      // 0 => no kmers_info
  // graph stage
  double   overlap_play = 1.3, errors = 3.0, density = 0.029, min_length = 100.0;
  bool     bases = false;
  int      tiling = 1;                         // 0 none 1 greedy 2 maximal 3 weighted
  int      trim = 0;                           // 0 none 1 match 2 branch
};

struct off_lis {
  std::vector<std::pair<int,int>> offsets;     // (pb, sr)
  std::vector<uint32_t>           lis;
};
struct mer_lists {
  uint32_t sr = 0;
  off_lis  fwd, bwd;
};

struct coords {
  int      rs, re, qs, qe, nb_mers;
  unsigned pb_cons, sr_cons, pb_cover, sr_cover;
  uint64_t rl, ql;
  bool     rn;
  uint32_t sr;
  bool     use_bwd_name;
  std::vector<int> kmers_info, bases_info;
  double   stretch, offset, avg_err;
  unsigned k;
  const std::vector<uint32_t>& unitigs(const sr_index& idx) const {
    return use_bwd_name ? idx.srs[sr].bwd_u : idx.srs[sr].fwd_u;
  }
  const std::string& name(const sr_index& idx) const {
    return use_bwd_name ? idx.srs[sr].bwd_name : idx.srs[sr].name;
  }
};

// chaining; lis_align.hpp:139-204 (window_size == 1 only)
std::vector<uint32_t> chain(const std::vector<std::pair<int,int>>& X, double a, double b, double C, uint32_t window = 1);

// coarse_aligner.cc:81-141; groups are returned in increasing super-read index
void fetch_super_reads(const sr_index& idx, const std::string& read, int max_count, std::vector<mer_lists>& groups);

// pb_aligner.cc:11-82
coords compute_coords_info(const sr_index& idx, const mer_lists& ml, uint64_t pb_size, const params& p,
                           const std::vector<int>* unitigs_lengths);

// same with the mer length given (the fine pass: align_k = --fine-mer, forward = true)
coords compute_coords_info_k(const sr_index& idx, const mer_lists& ml, uint64_t pb_size, const params& p,
                             const std::vector<int>* unitigs_lengths, unsigned k, bool forward);

// fine_aligner.hpp:49-58 + fine_aligner.cc:7-51: one window per coarse row, all hits of the shorter
// mer that fall on the row's super-read inside the window, accept-all chaining, coords with
// align_k = fine_k.  Output sorted like align_read's (ties: super-read index, then coarse order).
void fine_align_read(const sr_index& idx, const std::string& read, const params& p, unsigned fine_k,
                     const std::vector<int>* unitigs_lengths, const std::vector<coords>& coarse, std::vector<coords>& out);

// coarse_aligner.cc:42-60; output sorted by (rs, re, ql, sr) (create_mega_reads.cc:69-77 + tie rule)
void align_read(const sr_index& idx, const std::string& read, const params& p,
                const std::vector<int>* unitigs_lengths,
                std::vector<mer_lists>& groups, std::vector<coords>& out);

// overlap_graph.hpp:177-262 + overlap_graph.cc:7-299; appends the record text for one read
void mega_reads_for_read(const sr_index& idx, const std::vector<coords>& sorted_coords, const std::string& name,
                         uint64_t pb_size, const params& p, const std::vector<int>& unitigs_lengths,
                         const std::vector<std::string>* unitigs_sequences, std::string& out);

// jf_aligner.cc:41-70 (compact format)
void print_coords(const sr_index& idx, const std::vector<coords>& sorted_coords, const std::string& name,
                  uint64_t pb_size, std::string& out);

// compute_kmers_info, pb_aligner.cc:84-143
struct kmers_info_state {
  std::vector<int>& mers; std::vector<int>& bases;
  const std::vector<uint32_t>& u; unsigned cunitig = 0; int cend = 0; int prev_pos;
  unsigned k, unitigs_k; const std::vector<int>* ul; bool active;
  kmers_info_state(std::vector<int>& m, std::vector<int>& b, const std::vector<uint32_t>& u_,
                   unsigned unitigs_k_, unsigned k_, const std::vector<int>* ul_);
  void add_mer(int pos);
};

std::vector<uint32_t> parse_sr_name(const std::string& name);   // super_read_name.cc:74-90
int sr_overlap(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b);  // super_read_name.cc:49-72

} // namespace oport
#endif
