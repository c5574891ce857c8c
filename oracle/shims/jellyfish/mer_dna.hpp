// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for Jellyfish 2's
// <jellyfish/mer_dna.hpp>, which is not vendored in the reference and absent
// from this image (reference configure.ac:17 asks pkg-config for jellyfish-2.0).
// It provides exactly the API surface the reference hot path touches
// (jf_aligner.hpp:41-54, superread_parser.hpp:19-51,177-192,
// coarse_aligner.cc:8-15): a single-word (k <= 32) 2-bit k-mer, first base in
// the most significant position, A=0 C=1 G=2 T=3.
#ifndef ORACLE_SHIM_JELLYFISH_MER_DNA_HPP
#define ORACLE_SHIM_JELLYFISH_MER_DNA_HPP
#include <cstdint>
#include <string>
#include <ostream>
#include <stdexcept>

namespace jellyfish { namespace mer_dna_ns {

template<typename T> struct mer_base { };   // name only (superread_parser.hpp:21)

template<typename T, int CI>
class mer_base_static : public mer_base<T> {
  static unsigned int k_;
  T w_;
  static T mask() { return k_ >= 32 ? ~(T)0 : (((T)1 << (2 * k_)) - 1); }
public:
  struct base_proxy {
    int c;
    operator char() const { return "ACGT"[c & 3]; }
    int code() const { return c; }
  };

  mer_base_static() : w_(0) { }
  mer_base_static(const mer_base_static& o) : w_(o.w_) { }
  explicit mer_base_static(const std::string& s) : w_(0) { *this = s; }
  explicit mer_base_static(const char* s) : w_(0) { *this = std::string(s); }
  mer_base_static& operator=(const mer_base_static& o) { w_ = o.w_; return *this; }
  mer_base_static& operator=(const std::string& s) {
    if(s.size() < k_) throw std::length_error("mer_dna shim: string too short");
    w_ = 0;
    for(unsigned int i = 0; i < k_; ++i) {
      int c = code(s[i]);
      w_ = (w_ << 2) | (T)(c < 0 ? 0 : c);
    }
    return *this;
  }

  static unsigned int k() { return k_; }
  static unsigned int k(unsigned int n) { unsigned int o = k_; k_ = n; return o; }

  static int code(char c) {
    switch(c) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    default: return -1;
    }
  }
  static bool not_dna(int c) { return c < 0; }
  static int complement(int c) { return 3 - c; }

  // drop the first (most significant) base, append at the end
  int shift_left(int c) {
    int out = (int)((w_ >> (2 * (k_ - 1))) & 3);
    w_ = ((w_ << 2) | (T)(c & 3)) & mask();
    return out;
  }
  // drop the last (least significant) base, insert in front
  int shift_right(int c) {
    int out = (int)(w_ & 3);
    w_ = (w_ >> 2) | ((T)(c & 3) << (2 * (k_ - 1)));
    return out;
  }
  char shift_left(char c)  { int x = code(c); if(x < 0) return 'N'; return "ACGT"[shift_left(x)]; }
  char shift_right(char c) { int x = code(c); if(x < 0) return 'N'; return "ACGT"[shift_right(x)]; }

  // base(0) is the LAST character of the k-mer
  base_proxy base(unsigned int i) const { base_proxy p = { (int)((w_ >> (2 * i)) & 3) }; return p; }

  T get_bits(unsigned int start, unsigned int len) const {
    if(len == 0) return 0;
    T r = w_ >> start;
    return len >= 8 * sizeof(T) ? r : (r & (((T)1 << len) - 1));
  }
  T word(unsigned int) const { return w_; }

  mer_base_static get_reverse_complement() const {
    mer_base_static r;
    T x = w_;
    for(unsigned int i = 0; i < k_; ++i, x >>= 2)
      r.w_ = (r.w_ << 2) | (T)(3 - (x & 3));
    return r;
  }
  void reverse_complement() { *this = get_reverse_complement(); }
  void polyA() { w_ = 0; }

  bool operator<(const mer_base_static& o) const { return w_ < o.w_; }
  bool operator>(const mer_base_static& o) const { return w_ > o.w_; }
  bool operator==(const mer_base_static& o) const { return w_ == o.w_; }
  bool operator!=(const mer_base_static& o) const { return w_ != o.w_; }

  std::string to_str() const {
    std::string s(k_, 'A');
    for(unsigned int i = 0; i < k_; ++i) s[i] = "ACGT"[(w_ >> (2 * (k_ - 1 - i))) & 3];
    return s;
  }
};
template<typename T, int CI> unsigned int mer_base_static<T, CI>::k_ = 22;

template<typename T, int CI>
std::ostream& operator<<(std::ostream& os, const mer_base_static<T, CI>& m) { return os << m.to_str(); }

} // namespace mer_dna_ns
typedef mer_dna_ns::mer_base_static<uint64_t, 0> mer_dna;
} // namespace jellyfish
#endif
