// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT.  See oracle_port.hpp.
#include <limits>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include "oracle_port.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace oport {

// ===========================================================================
// super-read names  (reference: src_jf_aligner/super_read_name.cc:74-90)
// ===========================================================================
std::vector<uint32_t> parse_sr_name(const std::string& name) {
  std::vector<uint32_t> res;
  if(name.empty()) return res;
  try {
    size_t pn = 0;
    for(size_t n = name.find('_'); n != std::string::npos; pn = n + 1, n = name.find('_', pn)) {
      const uint32_t id = std::stoul(name.c_str() + pn);
      res.push_back(((id & 0x7fffffffu) << 1) | (name[n - 1] == 'R'));
    }
    const uint32_t id = std::stoul(name.c_str() + pn);
    res.push_back(((id & 0x7fffffffu) << 1) | (name[name.size() - 1] == 'R'));
  } catch(std::invalid_argument&) {
    res.clear();
  }
  return res;
}

static std::string unitigs_to_name(const std::vector<uint32_t>& u) {
  std::string s;
  for(size_t i = 0; i < u.size(); ++i) {
    if(i) s += '_';
    s += std::to_string(u[i] >> 1);
    s += (u[i] & 1) ? 'R' : 'F';
  }
  return s;
}

// dovetail overlap: largest t with last t of a == first t of b (super_read_name.cc:49-72)
int sr_overlap(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
  const uint32_t sa = a.size(), sb = b.size();
  if(sa < 2 || sb < 2) return 0;
  int32_t first = (int32_t)sa - (int32_t)sb + 1;
  if(first < 1) first = 1;
  for(uint32_t i = first; i < sa; ++i) {
    if(b[0] != a[i]) continue;
    uint32_t j = i + 1;
    while(j < sa && a[j] == b[j - i]) ++j;
    if(j == sa) return sa - i;
  }
  return 0;
}

// ===========================================================================
// index
// ===========================================================================
void sr_index::append_fasta(const std::string& path) {
  std::ifstream is(path);
  if(!is.good()) throw std::runtime_error("Can't open file " + path);
  if(is.peek() != '>') throw std::runtime_error("Not in fasta format");
  if(starts.empty()) starts.push_back(0);
  std::string header, line;
  int c = is.peek();
  while(c != EOF) {
    std::getline(is, header);
    const uint64_t old = text.size();
    for(c = is.peek(); c != '>' && c != EOF; c = is.peek()) {
      std::getline(is, line);
      for(unsigned char x : line) text.push_back(((x >> 1) ^ (x >> 2)) & 3);   // compact_dna.hpp:102-107
    }
    if(text.size() > old) {
      super_read sr;
      sr.name  = header.substr(1);
      sr.fwd_u = parse_sr_name(sr.name);
      if(!sr.fwd_u.empty()) {                       // frag_info.hpp:25-33
        sr.bwd_u.resize(sr.fwd_u.size());
        for(size_t i = 0; i < sr.fwd_u.size(); ++i) sr.bwd_u[i] = sr.fwd_u[sr.fwd_u.size() - 1 - i] ^ 1;
        sr.bwd_name = unitigs_to_name(sr.bwd_u);
      } else {
        sr.bwd_name = sr.name;
      }
      sr.start = old;
      sr.len   = text.size() - old;
      srs.push_back(sr);
      starts.push_back(text.size());
    }
  }
}

uint64_t sr_index::kmer_at(uint64_t pos, unsigned len) const {
  uint64_t v = 0;
  const uint64_t N = text.size();
  for(unsigned i = 0; i < len; ++i) v = (v << 2) | (pos + i < N ? text[pos + i] : 0);
  return v;
}

void sr_index::build(unsigned m_, unsigned k_) {
  m = m_; k = k_;
  if(!(m < k) || k > 31) throw std::runtime_error("oracle port requires psa-min < mer <= 31");
  const uint64_t N = text.size();
  if(N < m) throw std::runtime_error("text shorter than psa-min");
  const uint64_t nsa = N - m + 1;
  // order = (k-mer padded with A, position descending)   (mer_sa_imp.hpp:352-366)
  std::vector<std::pair<uint64_t, uint64_t>> keyed(nsa);
  uint64_t roll = kmer_at(0, k);
  const uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
  for(uint64_t p = 0; p < nsa; ++p) {
    keyed[p] = std::make_pair(roll, ~p);
    roll = ((roll << 2) | (p + k < N ? text[p + k] : 0)) & mask;
  }
  std::sort(keyed.begin(), keyed.end());
  sa.resize(nsa);
  counts.assign(((size_t)1 << (2 * m)) + 1, 0);
  for(uint64_t i = 0; i < nsa; ++i) {
    sa[i] = ~keyed[i].second;
    ++counts[(keyed[i].first >> (2 * (k - m))) + 1];
  }
  for(size_t i = 1; i < counts.size(); ++i) counts[i] += counts[i - 1];   // counts[i] = #entries with m-mer < i
}

void sr_index::search(uint64_t mer, uint64_t& index_out, uint64_t& nb_out) const {
  const uint64_t N = text.size();
  const uint64_t pre = mer >> (2 * (k - m));
  uint64_t lo = counts[pre], hi = counts[pre + 1];
  // lower bound on padded key
  uint64_t a = lo, b = hi;
  while(a < b) { const uint64_t mid = (a + b) / 2; if(kmer_at(sa[mid], k) < mer) a = mid + 1; else b = mid; }
  const uint64_t first = a;
  b = hi;
  while(a < b) { const uint64_t mid = (a + b) / 2; if(kmer_at(sa[mid], k) <= mer) a = mid + 1; else b = mid; }
  uint64_t last = a, f = first;
  while(f < last && sa[f] + k > N) ++f;     // tail-short suffixes never match (mer_sa_imp.hpp:399-406)
  nb_out    = last - f;
  index_out = nb_out ? f : 0;
}

void sr_index::search_k(uint64_t mer, unsigned kk, uint64_t& index_out, uint64_t& nb_out) const {
  // mer_sa_imp.hpp:369-381 (pattern not longer than psa-min: the counts table alone) and :382-479
  // (longer: search on the first kk bases); both return the SA entries whose text starts with the
  // pattern, entries with fewer than kk bases left never match.
  const uint64_t N = text.size();
  uint64_t a = 0, b = sa.size();
  while(a < b) { const uint64_t mid = (a + b) / 2; if(kmer_at(sa[mid], kk) < mer) a = mid + 1; else b = mid; }
  const uint64_t first = a;
  b = sa.size();
  while(a < b) { const uint64_t mid = (a + b) / 2; if(kmer_at(sa[mid], kk) <= mer) a = mid + 1; else b = mid; }
  uint64_t last = a, f = first;
  while(f < last && sa[f] + kk > N) ++f;
  nb_out    = last - f;
  index_out = nb_out ? f : 0;
}

bool sr_index::locate_k(uint64_t x, unsigned kk, uint32_t& sr, int32_t& off) const {
  const size_t i = std::upper_bound(starts.begin(), starts.end(), x) - starts.begin() - 1;
  if(x + kk > starts[i + 1]) return false;
  sr  = i;
  off = (int32_t)(x - starts[i] + 1);
  return true;
}

bool sr_index::locate(uint64_t x, uint32_t& sr, int32_t& off) const {
  const size_t i = std::upper_bound(starts.begin(), starts.end(), x) - starts.begin() - 1;
  if(x + k > starts[i + 1]) return false;
  sr  = i;
  off = (int32_t)(x - starts[i] + 1);
  return true;
}

// ===========================================================================
// chaining  (reference: src_lis/lis_align.hpp:139-204; the window of lis_align.hpp:17-45 as a ring per list element)
// ===========================================================================
// ORACLE_CHAIN_STATS=1: totals of the list walk (printed at exit), used to size the device kernels
namespace {
struct chain_stats_t {
  std::atomic<uint64_t> lists{0}, hits{0}, front{0}, steps{0}, none{0}, none_steps{0}, depth{0}, big_lists{0}, big_hits{0}, big_steps{0}, big_depth{0}, big_front{0};
  bool on = getenv("ORACLE_CHAIN_STATS") != nullptr;
  ~chain_stats_t() {
    if(on) fprintf(stderr, "chain stats: lists %llu hits %llu front-hit %llu steps %llu no-predecessor %llu (steps %llu) insert-depth %llu | lists>512: %llu hits %llu front %llu steps %llu depth %llu\n",
      (unsigned long long)lists, (unsigned long long)hits, (unsigned long long)front, (unsigned long long)steps, (unsigned long long)none, (unsigned long long)none_steps,
      (unsigned long long)depth, (unsigned long long)big_lists, (unsigned long long)big_hits, (unsigned long long)big_front, (unsigned long long)big_steps, (unsigned long long)big_depth);
  }
} chain_stats;
}

// window > 1 (lis_align.hpp:17-45,162-163): restated literally -- every list element carries the last `window` steps
// of its chain in a ring with a running sum; the mer predicate sees that sum plus the new step (minus the step that
// falls out of a full ring) and is only applied when the ring will be full with the new step.
static std::vector<uint32_t> chain_window(const std::vector<std::pair<int,int>>& X, double a, double b, double C, uint32_t W) {
  struct ring {
    std::vector<std::pair<double,double>> v; size_t next = 0; bool filled = false; double s1 = 0, s2 = 0;
    explicit ring(size_t n) : v(n, std::make_pair(0.0, 0.0)) { }
    bool will_be_filled() const { return filled || next == v.size() - 1; }
    std::pair<double,double> test_sum(double d1, double d2) const {
      double r1 = s1 + d1, r2 = s2 + d2;
      if(filled || next > 0) { r1 -= v[next].first; r2 -= v[next].second; }
      return std::make_pair(r1, r2);
    }
    void push(double d1, double d2) {
      const auto t = test_sum(d1, d2);
      s1 = t.first; s2 = t.second;
      v[next] = std::make_pair(d1, d2);
      next = (next + 1) % v.size();
      filled = filled || next == 0;
    }
  };
  struct elt { uint32_t idx, len; ring win; double span_pb, span_sr; };
  const uint32_t N = X.size();
  std::vector<elt>      L;
  std::vector<uint32_t> P(N, N);
  uint32_t longest = 0, best = 0;
  for(uint32_t i = 0; i < N; ++i) {
    elt e = { i, 1, ring(W), 0.0, 0.0 };
    int prev = -1;
    for(size_t p = 0; p < L.size(); ++p) {
      const uint32_t j = L[p].idx;
      if(X[i].second > X[j].second) {
        const double d1 = X[i].first - X[j].first, d2 = X[i].second - X[j].second;
        const auto ns = L[p].win.test_sum(d1, d2);
        const double t1 = a * ns.second, t2 = a * ns.first;
        if(!L[p].win.will_be_filled() || (ns.first <= b + t1 && ns.second <= b + t2 && ns.first <= C && ns.second <= C)) {
          e.len = L[p].len + 1;
          P[i]  = j;
          e.win = L[p].win;
          e.win.push(d1, d2);
          e.span_pb = L[p].span_pb + d1;
          e.span_sr = L[p].span_sr + d2;
          break;
        }
      }
      if(prev < 0 || L[p].len < L[prev].len) prev = p;
    }
    L.insert(L.begin() + (prev + 1), e);
    const double s1 = a * e.span_sr, s2 = a * e.span_pb;
    if(longest < e.len && e.span_pb <= s1 && e.span_sr <= s2) { longest = e.len; best = i; }
  }
  std::vector<uint32_t> res(longest);
  for(uint32_t t = 0, cur = best; t < longest; ++t, cur = P[cur]) res[longest - 1 - t] = cur;
  return res;
}

std::vector<uint32_t> chain(const std::vector<std::pair<int,int>>& X, double a, double b, double C, uint32_t window) {
  if(window != 1) return chain_window(X, a, b, C, window);
  struct elt { uint32_t idx, len; double span_pb, span_sr; };
  const uint32_t N = X.size();
  uint64_t st_front = 0, st_steps = 0, st_none = 0, st_none_steps = 0, st_depth = 0;
  std::vector<elt>      L;       // list order of the reference's forward_list
  std::vector<uint32_t> P(N, N);
  uint32_t longest = 0, best = 0;
  L.reserve(N);
  for(uint32_t i = 0; i < N; ++i) {
    elt e = { i, 1, 0.0, 0.0 };
    int prev = -1;
    size_t walked = L.size();
    for(size_t p = 0; p < L.size(); ++p) {      // every stored len is >= 1 == e.len: the walk only ends on a hit
      const uint32_t j = L[p].idx;
      if(X[i].second > X[j].second) {
        const double d1 = X[i].first - X[j].first, d2 = X[i].second - X[j].second;
        const double t1 = a * d2, t2 = a * d1;
        if(d1 <= b + t1 && d2 <= b + t2 && d1 <= C && d2 <= C) {
          e.len = L[p].len + 1;
          P[i]  = j;
          e.span_pb = L[p].span_pb + d1;
          e.span_sr = L[p].span_sr + d2;
          walked = p;
          break;
        }
      }
      if(prev < 0 || L[p].len < L[prev].len) prev = p;
    }
    if(chain_stats.on) {
      if(e.len > 1 && walked == 0) ++st_front;
      st_steps += walked;
      if(e.len == 1) { ++st_none; st_none_steps += walked; }
      st_depth += prev + 1;
    }
    L.insert(L.begin() + (prev + 1), e);
    const double s1 = a * e.span_sr, s2 = a * e.span_pb;
    if(longest < e.len && e.span_pb <= s1 && e.span_sr <= s2) { longest = e.len; best = i; }
  }
  if(chain_stats.on) {
    chain_stats.lists++; chain_stats.hits += N; chain_stats.front += st_front; chain_stats.steps += st_steps;
    chain_stats.none += st_none; chain_stats.none_steps += st_none_steps; chain_stats.depth += st_depth;
    if(N > 512) { chain_stats.big_lists++; chain_stats.big_hits += N; chain_stats.big_front += st_front; chain_stats.big_steps += st_steps; chain_stats.big_depth += st_depth; }
  }
  std::vector<uint32_t> res(longest);
  for(uint32_t t = 0, cur = best; t < longest; ++t, cur = P[cur]) res[longest - 1 - t] = cur;
  return res;
}

// lis_align::accept_all for both predicates (fine_aligner.cc:43-46): only the strict increase of the
// super-read offset decides whether an element extends a list entry
static std::vector<uint32_t> chain_accept_all(const std::vector<std::pair<int,int>>& X) {
  struct elt { uint32_t idx, len; };
  const uint32_t N = X.size();
  std::vector<elt>      L;
  std::vector<uint32_t> P(N, N);
  uint32_t longest = 0, best = 0;
  L.reserve(N);
  for(uint32_t i = 0; i < N; ++i) {
    elt e = { i, 1 };
    int prev = -1;
    for(size_t p = 0; p < L.size(); ++p) {
      if(X[i].second > X[L[p].idx].second) { e.len = L[p].len + 1; P[i] = L[p].idx; break; }
      if(prev < 0 || L[p].len < L[prev].len) prev = p;
    }
    L.insert(L.begin() + (prev + 1), e);
    if(longest < e.len) { longest = e.len; best = i; }
  }
  std::vector<uint32_t> res(longest);
  for(uint32_t t = 0, cur = best; t < longest; ++t, cur = P[cur]) res[longest - 1 - t] = cur;
  return res;
}

// ===========================================================================
// seed selection + hit expansion  (reference: coarse_aligner.cc:81-141)
// ===========================================================================
static bool is_ssr(uint64_t mer, unsigned k) {            // coarse_aligner.cc:8-15
  uint64_t r = mer;
  for(int i = 0; i < 2; ++i) {
    r = (r >> 2) | ((r & 3) << (2 * (k - 1)));
    if(r == mer) return true;
  }
  return false;
}

void fetch_super_reads(const sr_index& idx, const std::string& read, int max_count, std::vector<mer_lists>& groups) {
  const unsigned k = idx.k;
  const uint64_t mask = (1ULL << (2 * k)) - 1;
  struct list_info { uint64_t fi, fn, bi, bn; bool canonical; int offset; };
  std::vector<list_info> lists;
  std::vector<uint64_t>  sizes;
  uint64_t m = 0, rm = 0;
  unsigned len = 0;
  uint32_t flag = 1;
  for(size_t i = 0; i < read.size(); ++i) {               // jf_aligner.hpp:41-52,113-123
    int code;
    switch(read[i]) {
    case 'a': case 'A': code = 0; break;
    case 'c': case 'C': code = 1; break;
    case 'g': case 'G': code = 2; break;
    case 't': case 'T': code = 3; break;
    default: code = -1;
    }
    if(code < 0) { len = 0; continue; }
    ++len;
    m  = ((m << 2) | (uint64_t)code) & mask;
    rm = (rm >> 2) | ((uint64_t)(3 - code) << (2 * (k - 1)));
    if(len < k) continue;
    if(is_ssr(m, k)) continue;
    if(len <= 17) {
      flag = 1 - flag;
      if(flag == 1) continue;
    }
    const bool canonical = m < rm;
    list_info li;
    idx.search(canonical ? m : rm, li.fi, li.fn);
    idx.search(canonical ? rm : m, li.bi, li.bn);
    const uint64_t size = li.fn + li.bn;
    if(size == 0 || (max_count && size >= (uint64_t)max_count)) continue;
    li.canonical = canonical;
    li.offset    = (int)(i + 1) - (int)k + 1;
    lists.push_back(li);
    sizes.push_back(size);
  }

  // smallest t with #{size <= t} > round(0.99 * #lists)   (coarse_aligner.cc:117-125)
  const uint32_t sum_thresh = (uint32_t)std::round(lists.size() * 0.99);
  uint64_t threshold = (uint64_t)(max_count ? max_count : INT_MAX) + 1;
  if(sum_thresh < sizes.size()) {
    std::vector<uint64_t> sorted(sizes);
    std::nth_element(sorted.begin(), sorted.begin() + sum_thresh, sorted.end());
    threshold = sorted[sum_thresh];
  }

  std::map<uint32_t, mer_lists> by_sr;
  for(size_t l = 0; l < lists.size(); ++l) {
    if(sizes[l] > threshold) continue;
    const list_info& li = lists[l];
    for(int pass = 0; pass < 2; ++pass) {
      const uint64_t b = pass ? li.bi : li.fi, e = b + (pass ? li.bn : li.fn);
      for(uint64_t r = b; r < e; ++r) {
        uint32_t sr; int32_t off;
        if(!idx.locate(idx.sa[r], sr, off)) continue;
        if(pass) off = -off;
        if(!li.canonical) off = -off;
        mer_lists& ml = by_sr[sr];
        ml.sr = sr;
        (off > 0 ? ml.fwd : ml.bwd).offsets.push_back(std::make_pair(li.offset, off));
      }
    }
  }
  groups.clear();
  for(auto& it : by_sr) groups.push_back(std::move(it.second));
}

// ===========================================================================
// coords  (reference: pb_aligner.cc:11-143, pb_aligner.hpp:151-174, least_square_2d.hpp:47-67)
// ===========================================================================
static const uint32_t invalid_id = 0x7fffffffu;
static inline uint32_t uid(const std::vector<uint32_t>& u, size_t i) { return i < u.size() ? (u[i] >> 1) : invalid_id; }

kmers_info_state::kmers_info_state(std::vector<int>& m, std::vector<int>& b, const std::vector<uint32_t>& u_,
                                   unsigned unitigs_k_, unsigned k_, const std::vector<int>* ul_)
  : mers(m), bases(b), u(u_), prev_pos(-(int)k_), k(k_), unitigs_k(unitigs_k_), ul(ul_), active(unitigs_k_ != 0) {
  if(!active) return;
  const uint32_t id = uid(u, 0);
  if(id != invalid_id && id < ul->size()) {
    mers.assign(2 * u.size() - 1, 0);
    bases.assign(2 * u.size() - 1, 0);
    cend = (*ul)[id];
  } else {
    mers.clear(); bases.clear();
  }
}

void kmers_info_state::add_mer(int pos) {
  if(!active) return;
  // Once an error has emptied the vectors the reference keeps indexing into them (UB that
  // happens to be harmless there); we stop instead.
  if(mers.empty()) return;
  const int K = (int)k, UK = (int)unitigs_k;
  const int sr_pos    = std::abs(pos);
  const int new_bases = std::min(K, sr_pos - prev_pos);
  auto fail = [&]() { mers.clear(); bases.clear(); };
  while(sr_pos + K > cend + 1) {
    if(cend >= sr_pos) {
      if(cunitig >= u.size() - 1) return fail();
      const int nb = cend - std::max(sr_pos, prev_pos + K) + 1;
      bases[2 * cunitig]     += nb;
      bases[2 * cunitig + 1] += nb;
    }
    const uint32_t id = uid(u, ++cunitig);
    if(id == invalid_id || id >= ul->size()) return fail();
    cend += (*ul)[id] - UK + 1;
  }
  ++mers[2 * cunitig];
  bases[2 * cunitig] += new_bases;
  int cendi = cend;
  for(unsigned i = cunitig; i < u.size() - 1 && sr_pos + K > cendi - UK + 1; ++i) {
    const int full_mer = sr_pos + UK > cendi + 1;
    mers[2 * i + 1] += full_mer;
    mers[2 * i + 2] += full_mer;
    const int nb = std::min(new_bases, sr_pos + K - cendi + UK - 2);
    bases[2 * i + 1] += nb;
    bases[2 * i + 2] += nb;
    const uint32_t id = uid(u, i + 1);
    if(id != invalid_id && id < ul->size()) cendi += (*ul)[id] - UK + 1;
    else return fail();
  }
  prev_pos = sr_pos;
}

coords compute_coords_info(const sr_index& idx, const mer_lists& ml, uint64_t pb_size, const params& p,
                           const std::vector<int>* unitigs_lengths) {
  return compute_coords_info_k(idx, ml, pb_size, p, unitigs_lengths, idx.k, p.forward);
}

coords compute_coords_info_k(const sr_index& idx, const mer_lists& ml, uint64_t pb_size, const params& p0,
                             const std::vector<int>* unitigs_lengths, unsigned k, bool forward) {
  params p = p0;
  p.forward = forward;
  const size_t nf = ml.fwd.lis.size(), nb = ml.bwd.lis.size();
  const bool fwd_align = nf >= nb;
  coords c;
  c.nb_mers = fwd_align ? nf : nb;
  c.pb_cons = c.sr_cons = 0; c.pb_cover = c.sr_cover = k;
  c.rl = pb_size; c.ql = idx.srs[ml.sr].len; c.rn = false; c.sr = ml.sr;
  c.use_bwd_name = p.forward && !fwd_align;
  c.stretch = c.offset = c.avg_err = 0; c.k = k;
  c.rs = c.re = c.qs = c.qe = 0;
  if(c.nb_mers == 0) return c;
  const auto& offsets = fwd_align ? ml.fwd.offsets : ml.bwd.offsets;
  const auto& lis     = fwd_align ? ml.fwd.lis : ml.bwd.lis;
  kmers_info_state ki(c.kmers_info, c.bases_info, c.unitigs(idx), p.unitigs_k, k, unitigs_lengths);

  // online least squares, x = super-read offset, y = read offset (least_square_2d.hpp:47-67)
  double EX = 0, EY = 0, EXX = 0, EXY = 0, VX = 0, CXY = 0, NB = 0;
  long n = 0;
  auto lsq_add = [&](double x, double y) {
    ++n;
    const double dX = x - EX;  EX += dX / n;  const double ndX = x - EX;  VX += dX * ndX;
    const double dY = y - EY;  EY += dY / n;  const double ndY = y - EY;
    const double dXX = x * x - EXX;  EXX += dXX / n;
    const double dXY = x * y - EXY;  EXY += dXY / n;
    CXY += dX * ndY;
    NB  += dXY * ndX - dXX * ndY;
  };
  std::pair<int,int> prev = offsets[lis[0]];
  auto mer_pos = [&](int sr_off) { return fwd_align ? sr_off : (int)((int64_t)c.ql + sr_off - (int64_t)k + 2); };
  ki.add_mer(mer_pos(prev.second));
  lsq_add(prev.second, prev.first);
  for(size_t t = 1; t < lis.size(); ++t) {
    const std::pair<int,int> cur = offsets[lis[t]];
    const unsigned pb_diff = cur.first - prev.first;
    c.pb_cons  += pb_diff == 1;
    c.pb_cover += std::min(k, pb_diff);
    const unsigned sr_diff = cur.second - prev.second;
    c.sr_cons  += sr_diff == 1;
    c.sr_cover += std::min(k, sr_diff);
    ki.add_mer(mer_pos(cur.second));
    lsq_add(cur.second, cur.first);
    prev = cur;
  }
  if(n == 1) {
    c.stretch = 1.0; c.offset = EY - EX; c.avg_err = 0;
  } else {
    const double a = c.stretch = CXY / VX;
    const double b = c.offset  = NB / VX;
    double e = 0;
    for(uint32_t v : lis) {
      const double prod = a * offsets[v].second;
      e += std::abs(prod + b - offsets[v].first);
    }
    c.avg_err = e / n;
  }
  c.rs = offsets[lis.front()].first;
  c.re = offsets[lis.back()].first + k - 1;
  c.qs = offsets[lis.front()].second;
  c.qe = offsets[lis.back()].second;
  if(c.qs < 0) {                                         // pb_aligner.hpp:151-167
    if(p.forward) {
      c.qs = (int)((int64_t)c.ql + c.qs - (int64_t)k + 2);
      c.qe = (int)((int64_t)c.ql + c.qe + 1);
      c.rn = true;
      const double t = c.stretch * (double)(c.ql + 1);
      c.offset -= t - (double)k;
    } else {
      c.qs = -c.qs + (int)k - 1;
      c.qe = -c.qe;
      c.stretch = -c.stretch;
      c.offset += (double)(k - 1);
    }
  } else {
    c.qe += k - 1;
  }
  return c;
}

static double imp_s(const coords& c) { return std::max(1.0, std::min((double)c.rl, c.stretch + c.offset)); }
static double imp_e(const coords& c) { const double t = c.stretch * (double)c.ql; return std::max(1.0, std::min((double)c.rl, t + c.offset)); }
static int imp_len(const coords& c) { return (int)std::labs(std::lrint(imp_e(c) - imp_s(c))) + 1; }

static void discard_lis(off_lis& l) {                     // pb_aligner.hpp:47-61
  if(l.lis.empty()) return;
  std::vector<std::pair<int,int>> kept;
  size_t li = 0;
  for(size_t i = 0; i < l.offsets.size(); ++i) {
    if(li < l.lis.size() && l.lis[li] == i) ++li; else kept.push_back(l.offsets[i]);
  }
  l.offsets.swap(kept);
}

void align_read(const sr_index& idx, const std::string& read, const params& p,
                const std::vector<int>* unitigs_lengths,
                std::vector<mer_lists>& groups, std::vector<coords>& out) {
  if(p.window_size < 1) throw std::runtime_error("oracle port: window-size must be at least 1");
  fetch_super_reads(idx, read, p.max_count, groups);
  out.clear();
  for(auto& ml : groups) {                                // coarse_aligner.cc:42-60
    ml.fwd.lis = chain(ml.fwd.offsets, p.stretch_factor, p.stretch_constant, p.stretch_cap, p.window_size);
    ml.bwd.lis = chain(ml.bwd.offsets, p.stretch_factor, p.stretch_constant, p.stretch_cap, p.window_size);
    while(true) {
      coords c = compute_coords_info(idx, ml, read.size(), p, unitigs_lengths);
      if(c.nb_mers == 0) break;
      if(std::fabs(c.stretch) == 0.0) break;
      if(p.matching_mers != 0.0 &&
         !(p.matching_mers * (double)(unsigned)((unsigned)imp_len(c) - idx.k + 1) <= (double)c.nb_mers)) break;
      if(p.matching_bases > 0.0 &&
         !(p.matching_bases * (double)(imp_len(c) - 2 * (int)idx.k) <= (double)c.pb_cover)) break;
      out.push_back(c);
      if(!p.max_match) break;
      off_lis& l = ml.fwd.lis.size() > ml.bwd.lis.size() ? ml.fwd : ml.bwd;   // pb_aligner.hpp:87-92
      discard_lis(l);
      l.lis = chain(l.offsets, p.stretch_factor, p.stretch_constant, p.stretch_cap, p.window_size);
    }
  }
  // create_mega_reads.cc:69-77 sorts (unstably) by (rs, re, ql); ties are broken here by
  // emission order, i.e. by super-read index -- the canonical order of this project.
  if(getenv("ORACLE_TIE_REVERSE")) std::reverse(out.begin(), out.end());   // diagnostic: break (rs, re, ql) ties the other way
  std::stable_sort(out.begin(), out.end(), [](const coords& a, const coords& b) {
    return a.rs < b.rs || (a.rs == b.rs && (a.re < b.re || (a.re == b.re && a.ql < b.ql)));
  });
}

// ===========================================================================
// fine pass (reference: fine_aligner.hpp:49-58, fine_aligner.cc:7-51)
// ===========================================================================
void fine_align_read(const sr_index& idx, const std::string& read, const params& p, unsigned fine_k,
                     const std::vector<int>* unitigs_lengths, const std::vector<coords>& coarse, std::vector<coords>& out) {
  struct window { double begin, end; mer_lists ml; };
  // the reference keys this map by the address of the super-read's name (iteration order =
  // allocation order of the frag_info objects); only the order of exact (rs, re, ql) ties depends on it
  std::map<uint32_t, std::vector<window>> wins;
  for(const coords& c : coarse) {                          // prime_frags_pos, fine_aligner.hpp:49-58
    window w;
    const double s1 = c.stretch + c.offset;
    w.begin = std::max(0.0, s1 - c.avg_err);
    const double t = c.stretch * (double)c.ql;
    const double e1 = t + c.offset;
    const double e2 = e1 + c.avg_err;
    w.end = std::min((double)c.rl, e2 - (double)fine_k);
    w.ml.sr = c.sr;
    wins[c.sr].push_back(std::move(w));
  }
  // fetch_local_super_reads, fine_aligner.cc:7-36; the mer stream is parser_base::next (jf_aligner.hpp:113-123)
  const uint64_t mask = fine_k < 32 ? (1ULL << (2 * fine_k)) - 1 : ~0ULL;
  uint64_t m = 0, rm = 0;
  unsigned len = 0;
  for(size_t i = 0; i < read.size(); ++i) {
    int code;
    switch(read[i]) {
    case 'a': case 'A': code = 0; break;
    case 'c': case 'C': code = 1; break;
    case 'g': case 'G': code = 2; break;
    case 't': case 'T': code = 3; break;
    default: code = -1;
    }
    if(code < 0) { len = 0; continue; }
    ++len;
    m  = ((m << 2) | (uint64_t)code) & mask;
    rm = (rm >> 2) | ((uint64_t)(3 - code) << (2 * (fine_k - 1)));
    if(len < fine_k) continue;
    const bool canonical = m < rm;
    const int pb_off = (int)(i + 1) - (int)fine_k + 1;
    uint64_t ri[2], rn[2];
    idx.search_k(canonical ? m : rm, fine_k, ri[0], rn[0]);
    idx.search_k(canonical ? rm : m, fine_k, ri[1], rn[1]);
    for(int pass = 0; pass < 2; ++pass) {
      for(uint64_t r = ri[pass]; r < ri[pass] + rn[pass]; ++r) {
        uint32_t sr; int32_t off;
        if(!idx.locate_k(idx.sa[r], fine_k, sr, off)) continue;
        if(pass) off = -off;
        if(!canonical) off = -off;
        auto it = wins.find(sr);
        if(it == wins.end()) continue;
        for(window& w : it->second)
          if((double)pb_off >= w.begin && (double)pb_off <= w.end)
            (off > 0 ? w.ml.fwd : w.ml.bwd).offsets.push_back(std::make_pair(pb_off, off));
      }
    }
  }
  out.clear();
  const double inf = std::numeric_limits<double>::infinity();
  (void)inf;
  for(auto& it : wins) {
    for(window& w : it.second) {                           // fine_aligner.cc:43-50: accept-all LIS, every window gives a row
      w.ml.fwd.lis = chain_accept_all(w.ml.fwd.offsets);
      w.ml.bwd.lis = chain_accept_all(w.ml.bwd.offsets);
      out.push_back(compute_coords_info_k(idx, w.ml, read.size(), p, unitigs_lengths, fine_k, true));
    }
  }
  std::stable_sort(out.begin(), out.end(), [](const coords& a, const coords& b) {
    return a.rs < b.rs || (a.rs == b.rs && (a.re < b.re || (a.re == b.re && a.ql < b.ql)));
  });
}

// ===========================================================================
// overlap graph, tiling, printing (reference: overlap_graph.hpp:24-34,177-262; overlap_graph.cc:7-299)
// ===========================================================================
namespace {
struct node {
  bool   start_node, end_node;
  double imp_s, imp_e;
  int    parent, rank;
  int    lstart, lprev, lpath, lunitigs;
};
int find_root(std::vector<node>& nodes, int s) {          // union_find.cc:20-24
  if(nodes[s].parent != s) nodes[s].parent = find_root(nodes, nodes[s].parent);
  return nodes[s].parent;
}
void union_sets(std::vector<node>& nodes, int s1, int s2) {   // union_find.cc:6-18
  const int r1 = find_root(nodes, s1), r2 = find_root(nodes, s2);
  if(nodes[r1].rank > nodes[r2].rank) nodes[r2].parent = r1;
  else if(nodes[r1].rank < nodes[r2].rank) nodes[r1].parent = r2;
  else if(r1 != r2) { nodes[r2].parent = r1; ++nodes[r1].rank; }
}
struct mega_read {
  int start_node, end_node, start_unitig, end_unitig, start_offset, end_offset, nb_unitigs;
  double imp_s, imp_e, tiling_start, tiling_end, density;
};
struct interval { double lo, up; };
}

void mega_reads_for_read(const sr_index& idx, const std::vector<coords>& cs, const std::string& name,
                         uint64_t pb_size, const params& p, const std::vector<int>& ul,
                         const std::vector<std::string>* useq, std::string& out) {
  const int n = cs.size();
  const double K = p.unitigs_k;
  std::vector<node> nodes(n);
  std::vector<int>  order(n);
  for(int i = 0; i < n; ++i) {
    order[i] = i;
    node& nd = nodes[i];
    nd.start_node = nd.end_node = true;
    nd.imp_s = cs[i].stretch + cs[i].offset;
    { const double t = cs[i].stretch * (double)cs[i].ql; nd.imp_e = t + cs[i].offset; }
    nd.parent = i; nd.rank = 0; nd.lstart = nd.lprev = -1;
    nd.lpath = p.bases ? (int)cs[i].sr_cover : cs[i].nb_mers;
    nd.lunitigs = cs[i].unitigs(idx).size();
  }
  std::sort(order.begin(), order.end(), [&](int i, int j) {
    return nodes[i].imp_s < nodes[j].imp_s || (nodes[i].imp_s == nodes[j].imp_s && nodes[i].imp_e < nodes[j].imp_e);
  });

  // traverse (overlap_graph.cc:7-59)
  for(int a = 0; a < n; ++a) {
    const int ii = order[a];
    node& ni = nodes[ii];
    const coords& ci = cs[ii];
    if(ni.imp_e >= (double)ci.rl) continue;
    const auto& ui = ci.unitigs(idx);
    for(int b = a + 1; b < n; ++b) {
      const int jj = order[b];
      node& nj = nodes[jj];
      const coords& cj = cs[jj];
      if(nj.imp_s <= 1) continue;
      if(ni.imp_e > nj.imp_e + 31) continue;
      const double position_len = ni.imp_e - nj.imp_s;
      const double error1 = ci.avg_err + cj.avg_err;
      const double error  = p.errors * error1;
      { const double t = position_len * p.overlap_play; if(t + error < K) break; }
      const auto& uj = cj.unitigs(idx);
      const int nb_u = sr_overlap(ui, uj);
      if(!nb_u) continue;
      if(ui == uj) continue;
      int u_overlap_len = 0, common = 0;
      const std::vector<int>& info = p.bases ? cj.bases_info : cj.kmers_info;
      for(int u = 0; u < nb_u; ++u) {
        u_overlap_len += ul[uj[u] >> 1];
        common += info[2 * u];
        if(u > 0) common -= info[2 * u - 1];
      }
      u_overlap_len -= (nb_u - 1) * ((int)p.unitigs_k - 1);
      { const double t1 = p.overlap_play * position_len, t2 = p.overlap_play * ((double)u_overlap_len + error);
        if((double)u_overlap_len > t1 + error || position_len > t2) continue; }
      ni.end_node = false;
      nj.start_node = false;
      union_sets(nodes, ii, jj);
      const int nlpath = ni.lpath + (p.bases ? (int)cj.sr_cover : cj.nb_mers) - common;
      const node& si = ni.lstart == -1 ? ni : nodes[ni.lstart];
      const node& sj = nj.lstart == -1 ? nj : nodes[nj.lstart];
      if(nlpath > nj.lpath || (nlpath == nj.lpath && (nj.lstart == -1 || si.imp_s > sj.imp_s))) {
        nj.lpath    = nlpath;
        nj.lstart   = ni.lstart == -1 ? ii : ni.lstart;
        nj.lprev    = ii;
        nj.lunitigs = ni.lunitigs + (int)uj.size() - nb_u;
      }
    }
  }

  // best terminal node per component (overlap_graph.cc:61-161)
  std::map<int, mega_read> comps;
  for(int i = 0; i < n; ++i) {
    mega_read mr;
    mr.start_node = nodes[i].lstart == -1 ? i : nodes[i].lstart;
    mr.end_node = i;
    mr.start_unitig = 0;
    mr.nb_unitigs = nodes[i].lunitigs;
    mr.end_unitig = cs[i].kmers_info.size() / 2;
    mr.imp_s = cs[mr.start_node].stretch + cs[mr.start_node].offset;
    { const double t = cs[i].stretch * (double)cs[i].ql; mr.imp_e = t + cs[i].offset; }
    mr.tiling_start = cs[mr.start_node].rs;
    mr.tiling_end = cs[i].re;
    mr.start_offset = mr.end_offset = 0;
    if(p.trim != 0) {                                      // trim_match, overlap_graph.cc:78-114
      if(nodes[mr.start_node].imp_s < 1) {
        const coords& c = cs[mr.start_node];
        const auto& cu = c.unitigs(idx);
        int offset = 0;
        for(mr.start_unitig = 0; mr.start_unitig < (int)c.kmers_info.size(); mr.start_unitig += 2) {
          if(c.kmers_info[mr.start_unitig]) break;
          offset += ul[uid(cu, mr.start_unitig / 2)];
        }
        mr.start_unitig /= 2;
        mr.nb_unitigs -= mr.start_unitig;
        offset -= ((int)p.unitigs_k - 1) * mr.start_unitig;
        mr.start_offset = offset;
        { const double t = c.stretch * (double)(offset + 1); mr.imp_s = t + c.offset; }
      }
      const coords& c = cs[mr.end_node];
      if(nodes[mr.end_node].imp_e > (double)c.ql) {
        const auto& cu = c.unitigs(idx);
        int offset = 0;
        for(mr.end_unitig = (int)c.kmers_info.size() - 1; mr.end_unitig >= 0; mr.end_unitig -= 2) {
          if(c.kmers_info[mr.end_unitig]) break;
          offset += ul[uid(cu, mr.end_unitig / 2)];
        }
        mr.end_unitig /= 2;
        const int removed = (int)(c.kmers_info.size() / 2) - mr.end_unitig;
        mr.nb_unitigs -= removed;
        offset -= ((int)p.unitigs_k - 1) * removed;
        mr.end_offset = offset;
        { const double t = c.stretch * (double)(int64_t)((int64_t)c.ql - offset); mr.imp_e = t + c.offset; }
      }
    }
    const double len = std::min((double)pb_size + 0.5, mr.tiling_end) - std::max(0.5, mr.tiling_start);
    mr.density = (double)nodes[i].lpath / len;
    if(!nodes[i].end_node || mr.density < p.density || (mr.tiling_end - mr.tiling_start) < p.min_length) continue;
    const int root = find_root(nodes, i);
    auto it = comps.find(root);
    if(it == comps.end()) comps.insert(std::make_pair(root, mr));
    else {
      const node& o = nodes[it->second.end_node];
      if(nodes[i].lpath > o.lpath || (nodes[i].lpath == o.lpath && mr.density > it->second.density)) it->second = mr;
    }
  }
  if(comps.empty()) return;

  std::vector<mega_read> mrs;
  std::vector<int> sort_tiling, tiled;
  for(const auto& c : comps) { sort_tiling.push_back(mrs.size()); mrs.push_back(c.second); }

  auto by_pos = [&](int i, int j) {
    return mrs[i].imp_s < mrs[j].imp_s || (mrs[i].imp_s == mrs[j].imp_s && mrs[i].imp_e < mrs[j].imp_e);
  };
  auto greedy = [&]() {                                    // overlap_graph.cc:165-197
    std::vector<interval> covered, placed;               // covered: disjoint, sorted, touching merged
    for(const int it : sort_tiling) {
      const mega_read& mr = mrs[it];
      const interval pos = { mr.tiling_start, mr.tiling_end };
      const double plen = pos.lo < pos.up ? pos.up - pos.lo : 0.0;
      const double max_overlap = std::max(K * p.overlap_play, plen * (p.overlap_play - 0.9));
      bool large = false;
      for(const auto& c : covered) {
        const double lo = std::max(c.lo, pos.lo), up = std::min(c.up, pos.up);
        if(lo < up && up - lo >= max_overlap) { large = true; break; }
      }
      if(large) continue;
      bool contained = false;
      for(const auto& c : placed)
        if(!(pos.lo < pos.up) || (c.lo <= pos.lo && pos.up <= c.up)) { contained = true; break; }
      if(contained) continue;
      if(pos.lo < pos.up) {
        interval nw = pos;
        std::vector<interval> nv;
        bool put = false;
        for(const auto& c : covered) {
          if(c.up < nw.lo) nv.push_back(c);
          else if(nw.up < c.lo) { if(!put) { nv.push_back(nw); put = true; } nv.push_back(c); }
          else { nw.lo = std::min(nw.lo, c.lo); nw.up = std::max(nw.up, c.up); }
        }
        if(!put) nv.push_back(nw);
        covered.swap(nv);
      }
      placed.push_back(pos);
      tiled.push_back(it);
    }
  };
  switch(p.tiling) {
  case 1:                                                  // overlap_graph.hpp:211-221
    std::sort(sort_tiling.begin(), sort_tiling.end(),
              [&](int i, int j) { return nodes[mrs[j].end_node].lpath < nodes[mrs[i].end_node].lpath; });
    greedy();
    std::sort(tiled.begin(), tiled.end(), by_pos);
    break;
  case 3: {                                                // overlap_graph.hpp:223-239
    std::vector<double> w(mrs.size());
    for(const int i : sort_tiling) {
      const double d2 = mrs[i].density * mrs[i].density;
      w[i] = d2 * (double)(cs[mrs[i].end_node].re - cs[mrs[i].start_node].rs + 1);
    }
    std::sort(sort_tiling.begin(), sort_tiling.end(), [&](int i, int j) { return w[j] < w[i]; });
    greedy();
    std::sort(tiled.begin(), tiled.end(), by_pos);
    break;
  }
  case 2: {                                                // overlap_graph.hpp:241-250, overlap_graph.cc:212-252
    std::sort(sort_tiling.begin(), sort_tiling.end(),
              [&](int i, int j) { return mrs[i].tiling_end < mrs[j].tiling_end; });
    struct tinfo { int score; double pos; int node, previous, length; };
    std::vector<tinfo> info;
    auto it = sort_tiling.begin();
    info.push_back({ nodes[mrs[*it].end_node].lpath, mrs[*it].tiling_end, *it, -1, 1 });
    for(++it; it != sort_tiling.end(); ++it) {
      const double lstart = mrs[*it].tiling_start;
      const double key = std::min(lstart + K * p.overlap_play, mrs[*it].tiling_end);
      int i = (int)(std::upper_bound(info.begin(), info.end(), key,
                                     [](double x, const tinfo& y) { return x < y.pos; }) - info.begin()) - 1;
      while(i >= 0 && mrs[info[i].node].tiling_start >= lstart) i = info[i].previous;
      const int nscore = (i >= 0 ? info[i].score : 0) + nodes[mrs[*it].end_node].lpath;
      if(nscore > info.back().score)
        info.push_back({ nscore, mrs[*it].tiling_end, *it, i, (i >= 0 ? info[i].length : 0) + 1 });
    }
    tiled.resize(info.back().length);
    int ptr = info.size() - 1;
    for(auto r = tiled.rbegin(); r != tiled.rend(); ++r) { *r = info[ptr].node; ptr = info[ptr].previous; }
    std::sort(tiled.begin(), tiled.end(), by_pos);
    break;
  }
  default: break;
  }

  // print (overlap_graph.hpp:253-262, overlap_graph.cc:254-299)
  out += '>'; out += name; out += '\n';
  const std::vector<int>& final_order = tiled.empty() ? sort_tiling : tiled;
  char buf[256];
  for(const int cmr : final_order) {
    const mega_read& mr = mrs[cmr];
    const node& end_n = nodes[mr.end_node];
    const coords& end_c = cs[mr.end_node];
    const coords& start_c = cs[mr.start_node];
    std::vector<uint32_t> path(std::max(0, end_n.lunitigs), 0);
    auto prepend = [&](size_t offset, const std::vector<uint32_t>& rhs, size_t first, size_t last) -> size_t {
      if(first > last || first >= rhs.size()) return offset;
      const size_t to_copy = std::min(last, rhs.size() - 1) - first + 1;
      if(to_copy > offset) return offset;
      std::copy_n(rhs.begin() + first, to_copy, path.begin() + (offset - to_copy));
      return offset - to_copy;
    };
    const auto& eu = end_c.unitigs(idx);
    size_t offset = prepend(path.size(), eu, 0, eu.size() - 1);
    int node_j = mr.end_node, node_i = end_n.lprev;
    while(node_i >= 0) {
      const auto& iu = cs[node_i].unitigs(idx);
      const size_t overlap = (size_t)nodes[node_i].lunitigs + cs[node_j].unitigs(idx).size() - (size_t)nodes[node_j].lunitigs;
      const size_t end = iu.size() - 1 - overlap;
      offset = prepend(offset, iu, 0, end);
      node_j = node_i;
      node_i = nodes[node_i].lprev;
    }
    int sr_len = 0;
    for(int i = mr.start_unitig; i < mr.start_unitig + mr.nb_unitigs; ++i) sr_len += ul[uid(path, i)];
    sr_len -= (mr.nb_unitigs - 1) * ((int)p.unitigs_k - 1);
    const uint64_t qe_out = (uint64_t)(int64_t)(sr_len + mr.end_offset) - (end_c.ql - (uint64_t)(int64_t)end_c.qe);
    snprintf(buf, sizeof(buf), "%.2f %.2f %d %d %d %llu %d %.4f ", mr.imp_s, mr.imp_e, start_c.rs, end_c.re,
             start_c.qs - mr.start_offset, (unsigned long long)qe_out, end_n.lpath, mr.density);
    out += buf;
    out += unitigs_to_name(path);
    snprintf(buf, sizeof(buf), " %d", sr_len);
    out += buf;
    if(useq) {
      out += ' ';
      const size_t b = std::min((size_t)mr.start_unitig, path.size());
      const size_t e = std::min((size_t)(mr.start_unitig + mr.nb_unitigs), path.size());
      for(size_t i = b; i < e; ++i) {
        const std::string& s = useq->at(path[i] >> 1);
        const size_t skip = i == b ? 0 : (size_t)p.unitigs_k - 1;
        if(skip >= s.size()) continue;
        if(path[i] & 1) {
          for(size_t t = skip; t < s.size(); ++t) {
            char ch;
            switch(s[s.size() - 1 - t]) {
            case 'a': case 'A': ch = 'T'; break;
            case 'c': case 'C': ch = 'G'; break;
            case 'g': case 'G': ch = 'C'; break;
            case 't': case 'T': ch = 'A'; break;
            default: ch = 'N';
            }
            out += ch;
          }
        } else {
          out.append(s, skip, std::string::npos);
        }
      }
    }
    out += '\n';
  }
}

// "%g"-style default ostream formatting of doubles (6 significant digits), jf_aligner.cc:53-67
void print_coords(const sr_index& idx, const std::vector<coords>& cs, const std::string& name,
                  uint64_t pb_size, std::string& out) {
  if(cs.empty()) return;
  char buf[512];
  snprintf(buf, sizeof(buf), ">%zu %s\n", cs.size(), name.c_str());
  out += buf;
  for(const coords& c : cs) {
    snprintf(buf, sizeof(buf), "%d %d %d %d %d %u %u %u %u %llu %llu %g %g %g ", c.rs, c.re, c.qs, c.qe, c.nb_mers,
             c.pb_cons, c.sr_cons, c.pb_cover, c.sr_cover, (unsigned long long)pb_size, (unsigned long long)c.ql,
             c.stretch, c.offset, c.avg_err);
    out += buf;
    out += c.name(idx);
    for(size_t i = 0; i < c.kmers_info.size(); ++i) {
      snprintf(buf, sizeof(buf), " %d:%d", c.kmers_info[i], c.bases_info[i]);
      out += buf;
    }
    out += '\n';
  }
}

} // namespace oport
