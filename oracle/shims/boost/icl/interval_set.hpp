// TEST INFRASTRUCTURE ONLY (oracle build). Minimal stand-in for
// <boost/icl/interval_set.hpp> (Boost is absent from this image; reference
// configure.ac:20-21).  Implements only what overlap_graph.cc:163-190 uses:
// right_open_interval<double>, a *joining* interval_set (overlapping or
// touching intervals merge), set & interval, length(), contains().
#ifndef ORACLE_SHIM_BOOST_ICL_INTERVAL_SET_HPP
#define ORACLE_SHIM_BOOST_ICL_INTERVAL_SET_HPP
#include <algorithm>
#include <functional>
#include <vector>
namespace boost { namespace icl {

template<typename T>
class right_open_interval {
  T lo_, up_;
public:
  right_open_interval() : lo_(T()), up_(T()) { }
  right_open_interval(const T& lo, const T& up) : lo_(lo), up_(up) { }
  T lower() const { return lo_; }
  T upper() const { return up_; }
};

template<typename T> inline bool is_empty(const right_open_interval<T>& x) { return !(x.lower() < x.upper()); }
template<typename T> inline T length(const right_open_interval<T>& x) { return is_empty(x) ? T() : x.upper() - x.lower(); }
template<typename T>
inline bool contains(const right_open_interval<T>& super, const right_open_interval<T>& sub) {
  return is_empty(sub) || (super.lower() <= sub.lower() && sub.upper() <= super.upper());
}

template<typename T, template<class> class Compare = std::less, typename Interval = right_open_interval<T> >
class interval_set {
  std::vector<Interval> v_;   // disjoint, non touching, sorted by lower bound
public:
  typedef typename std::vector<Interval>::const_iterator const_iterator;
  typedef const_iterator iterator;
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  bool empty() const { return v_.empty(); }
  size_t iterative_size() const { return v_.size(); }

  interval_set& operator+=(const Interval& x) {
    if(is_empty(x)) return *this;
    T lo = x.lower(), up = x.upper();
    std::vector<Interval> nv;
    bool placed = false;
    for(const auto& y : v_) {
      if(y.upper() < lo) {                 // strictly before, not touching
        nv.push_back(y);
      } else if(up < y.lower()) {          // strictly after, not touching
        if(!placed) { nv.push_back(Interval(lo, up)); placed = true; }
        nv.push_back(y);
      } else {                             // overlap or touch: absorb
        lo = std::min(lo, y.lower());
        up = std::max(up, y.upper());
      }
    }
    if(!placed) nv.push_back(Interval(lo, up));
    v_.swap(nv);
    return *this;
  }
  // used only by operator& to collect pieces that are already disjoint
  void append_piece_(const Interval& x) { v_.push_back(x); }
};

template<typename T, template<class> class C, typename I>
inline interval_set<T, C, I> operator&(const interval_set<T, C, I>& s, const I& x) {
  interval_set<T, C, I> res;
  if(is_empty(x)) return res;
  for(const auto& y : s) {
    const T lo = std::max(y.lower(), x.lower());
    const T up = std::min(y.upper(), x.upper());
    if(lo < up) res.append_piece_(I(lo, up));
  }
  return res;
}
} }
#endif
