// TEST INFRASTRUCTURE ONLY (oracle build). Helpers shared by the hand-written
// stand-ins for the yaggo-generated *_cmdline.hpp headers (yaggo/ruby are
// absent; the generated headers are not in the reference, Makefile.am:148-153).
#ifndef ORACLE_SHIM_YAGGO_SHIM_HPP
#define ORACLE_SHIM_YAGGO_SHIM_HPP
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <getopt.h>
namespace yaggo_shim {
struct error_stream {
  std::ostringstream os_;
  const char* hint_;
  explicit error_stream(const char* hint) : hint_(hint) { }
  error_stream(const error_stream& o) : os_(o.os_.str()), hint_(o.hint_) { }
  template<typename T> error_stream& operator<<(const T& x) { os_ << x; return *this; }
  ~error_stream() {
    std::cerr << "Error: " << os_.str() << '\n' << hint_ << std::endl;
    std::exit(1);
  }
};
inline uint64_t conv_uint64(const char* s, bool suffix, bool& ok) {
  char* end = 0;
  ok = true;
  if(!s || !*s || *s == '-') { ok = false; return 0; }
  unsigned long long v = std::strtoull(s, &end, 0);
  if(end == s) { ok = false; return 0; }
  if(*end && suffix) {
    switch(*end) {
    case 'k': v *= 1000ULL; ++end; break;
    case 'M': v *= 1000000ULL; ++end; break;
    case 'G': v *= 1000000000ULL; ++end; break;
    case 'T': v *= 1000000000000ULL; ++end; break;
    case 'P': v *= 1000000000000000ULL; ++end; break;
    case 'E': v *= 1000000000000000000ULL; ++end; break;
    default: break;
    }
  }
  if(*end) ok = false;
  return v;
}
inline double conv_double(const char* s, bool& ok) {
  char* end = 0;
  ok = true;
  double v = std::strtod(s, &end);
  if(end == s || *end) ok = false;
  return v;
}
inline long conv_int(const char* s, bool& ok) {
  char* end = 0;
  ok = true;
  long v = std::strtol(s, &end, 0);
  if(end == s || *end) ok = false;
  return v;
}
}
#endif
