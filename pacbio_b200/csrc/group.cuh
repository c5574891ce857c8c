// Per-read grouping of the hits by super-read (group.cu).
#pragma once
#include "common.cuh"

struct group_sort_args {
  uint64_t* keys_in;  uint64_t* pays_in;      // hits in emission order (read-major); scratch afterwards
  uint64_t* keys_out; uint64_t* pays_out;     // every read's slice stably sorted by super-read
  uint8_t*  head;                             // head[i] = 1 where a (read, super-read) group starts
  const uint64_t* hit_off;                    // first hit of every tile (+ the total)
  const uint32_t* tile_first;                 // first tile of every read (+ the number of tiles)
  uint32_t nseq_all;                          // super-read index of the hits that belong to no super-read
  int      sr_bits;                           // nseq_all < 2^sr_bits
  unsigned long long* n_invalid_groups;       // counts the reads that have such hits (each is one group)
  uint32_t cap, idx_bits;                     // filled by launch_group_sort
};

// hits one CTA sorts in shared memory for an index whose super-read numbers take sr_bits bits (0: never)
uint32_t group_sort_capacity(int sr_bits);
int launch_read_hits_stats(mr_context* ctx, const uint64_t* hit_off, const uint32_t* tile_first, uint32_t nreads, uint32_t cap,
                           unsigned long long* big_hits);
int launch_group_sort(mr_context* ctx, group_sort_args A, uint32_t nreads);
