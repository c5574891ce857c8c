// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <jellyfish/err.hpp>:
// err::msg is an ostream-like builder convertible to std::string
// (reference call site: output_file.hpp:24).
#ifndef ORACLE_SHIM_JELLYFISH_ERR_HPP
#define ORACLE_SHIM_JELLYFISH_ERR_HPP
#include <sstream>
#include <string>
namespace jellyfish { namespace err {
class msg {
  std::ostringstream os_;
public:
  msg() { }
  template<typename T> msg& operator<<(const T& x) { os_ << x; return *this; }
  operator std::string() const { return os_.str(); }
};
} }
#endif
