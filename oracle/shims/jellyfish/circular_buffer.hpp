// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for Jellyfish's jellyfish/circular_buffer.hpp
// with the interface /root/reference/include/jflib/pool.hpp:59-61,145,183-230 uses: a bounded FIFO
// of integers with a `guard` value meaning "nothing", non-blocking enqueue / dequeue, close().
// A mutex replaces the lock-free ring of the original; behaviour seen by pool.hpp is the same.
#ifndef ORACLE_SHIM_CIRCULAR_BUFFER_HPP
#define ORACLE_SHIM_CIRCULAR_BUFFER_HPP
#include <atomic>
#include <cstddef>
#include <deque>
#include <limits>
#include <mutex>
namespace jflib {
template<typename T> inline T a_load(const T& x) { return __atomic_load_n(&x, __ATOMIC_SEQ_CST); }
template<typename T, typename U> inline void a_store(T& x, const U& v) { __atomic_store_n(&x, (T)v, __ATOMIC_SEQ_CST); }
template<typename T> inline T* a_load_ptr(T*& x) { return __atomic_load_n(&x, __ATOMIC_SEQ_CST); }
template<typename T, typename U> inline void a_store_ptr(T*& x, U* v) { __atomic_store_n(&x, (T*)v, __ATOMIC_SEQ_CST); }

template<typename T>
class circular_buffer {
  std::deque<T> q_;
  mutable std::mutex m_;
  size_t cap_;
  bool closed_;
public:
  static const T guard = std::numeric_limits<T>::max();
  explicit circular_buffer(size_t size) : cap_(size), closed_(false) { }
  bool enqueue(const T& v) {
    std::lock_guard<std::mutex> l(m_);
    if(q_.size() >= cap_) return false;
    q_.push_back(v);
    return true;
  }
  T dequeue() {
    std::lock_guard<std::mutex> l(m_);
    if(q_.empty()) return guard;
    const T v = q_.front();
    q_.pop_front();
    return v;
  }
  void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; }
  bool is_closed() const { std::lock_guard<std::mutex> l(m_); return closed_; }
};
template<typename T> const T circular_buffer<T>::guard;
}
#endif
