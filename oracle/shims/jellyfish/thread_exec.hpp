// TEST INFRASTRUCTURE ONLY (oracle build). <jellyfish/thread_exec.hpp> is
// included by pb_aligner.hpp:16 but nothing from it is used on the hot path.
#ifndef ORACLE_SHIM_JELLYFISH_THREAD_EXEC_HPP
#define ORACLE_SHIM_JELLYFISH_THREAD_EXEC_HPP
#endif
