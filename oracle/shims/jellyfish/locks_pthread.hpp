// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for Jellyfish's jellyfish/locks_pthread.hpp:
// the condition variable /root/reference/include/jflib/pool.hpp:48,186-241 waits on.
#ifndef ORACLE_SHIM_LOCKS_PTHREAD_HPP
#define ORACLE_SHIM_LOCKS_PTHREAD_HPP
#include <pthread.h>
#include <ctime>
namespace jellyfish { namespace locks {
class cond {
  pthread_mutex_t m_;
  pthread_cond_t  c_;
public:
  cond() { pthread_mutex_init(&m_, 0); pthread_cond_init(&c_, 0); }
  ~cond() { pthread_cond_destroy(&c_); pthread_mutex_destroy(&m_); }
  void lock() { pthread_mutex_lock(&m_); }
  void unlock() { pthread_mutex_unlock(&m_); }
  void wait() { pthread_cond_wait(&c_, &m_); }
  int  timedwait(time_t seconds) {
    struct timespec t;
    clock_gettime(CLOCK_REALTIME, &t);
    t.tv_sec += seconds;
    return pthread_cond_timedwait(&c_, &m_, &t);
  }
  void signal() { pthread_cond_signal(&c_); }
  void broadcast() { pthread_cond_broadcast(&c_); }
};
} }
#endif
