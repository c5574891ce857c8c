// Grouping of a batch's hits by (read, super-read): frags_pos_type of coarse_aligner.hpp:14 /
// coarse_aligner.cc:132-138, where the reference appends every hit to an unordered_map of per-super-read
// vectors while it walks one read.
//
// The expansion emits the hits read-major, in the order the reference visits them (read position, forward
// list before reverse list, suffix-array order), so the hits of one read are one contiguous slice of the hit
// arrays and grouping is a STABLE sort of that slice on the super-read index alone.  A read of the usual
// kind has a few thousand hits: the whole slice fits the shared memory of one SM.  One CTA per read therefore
//   * pulls the slice's payloads into shared memory with one bulk asynchronous copy (TMA engine, completes on
//     an mbarrier while the CTA is busy with the keys),
//   * turns every key into (super-read << idx_bits | position in the slice), a 32-bit word,
//   * sorts those words with an LSD radix sort that never leaves shared memory (7-bit digits; every warp ranks
//     a contiguous piece 32 keys at a time with match_any, so the order among equal digits is the input order),
//   * writes keys and payloads back in sorted order, fully coalesced, together with one byte per hit that
//     says whether a new (read, super-read) group starts there.
// Per hit that is 16 bytes read and 17 written, against (8 + 32) bytes per pass of a device-wide radix sort
// plus 16 for the group heads: 33 instead of 96 bytes on the yeast-size index (14-bit super-read index).
//
// A read whose hits do not fit (repeats; or an index with more than 2^18 super-reads, where idx_bits shrinks)
// is sorted by the same CTA out of global memory: LSD passes over its slice, 4096 hits at a time, the running
// digit offsets in shared memory.  The host looks at how many hits such reads hold (read_hits_stats_kernel)
// and keeps the device-wide radix sort for batches where they are the rule (the human-size shape).
//
// Hits whose k-mer straddles two super-reads carry the super-read index nseq_all: they sort to the end of
// their read's slice and form a group of their own, which the chaining kernels skip (chain.cu,
// classify_groups_kernel).
#include "align.cuh"
#include "group.cuh"

namespace {

constexpr int kGsRounds  = 16;              // keys per thread of the in-shared-memory sort
constexpr int kGsDigit   = 7;
constexpr int kGsRadix   = 1 << kGsDigit;
constexpr int kFbRounds  = 4;               // keys per thread and chunk of the global-memory fallback

template<int kThreads>
struct gs_layout {
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kCap   = kThreads * kGsRounds;
  static constexpr size_t pay_off = 0;                                             // uint64[kCap + 2], 16-byte aligned
  static constexpr size_t arr_off = pay_off + (size_t)(kCap + 2) * 8;              // uint32[kCap]
  static constexpr size_t cnt_off = arr_off + (size_t)kCap * 4;                    // uint32[kWarps][kGsRadix]
  static constexpr size_t tot_off = cnt_off + (size_t)kWarps * kGsRadix * 4;       // uint32[kGsRadix]
  static constexpr size_t dbs_off = tot_off + kGsRadix * 4;                        // uint32[kGsRadix]
  static constexpr size_t wsm_off = dbs_off + kGsRadix * 4;                        // uint32[4]
  static constexpr size_t bar_off = wsm_off + 16;                                  // uint64
  static constexpr size_t bytes   = bar_off + 8;
};

// rank of this lane's key among the keys of its warp's piece that have the same digit and come before it
// (earlier rounds through cnt, lower lanes of this round through the match mask); d == kGsRadix: no key
__device__ __forceinline__ uint32_t rank_round(uint32_t d, uint32_t* cnt_w, unsigned lane, unsigned lt) {
  const unsigned peers = __match_any_sync(MR_FULL_MASK, d);
  const unsigned before = __popc(peers & lt);
  uint32_t b = 0;
  if(d < (uint32_t)kGsRadix) b = cnt_w[d];
  __syncwarp();
  if(before == 0 && d < (uint32_t)kGsRadix) cnt_w[d] = b + __popc(peers);
  __syncwarp();
  return b + before;
}

// cnt[w][d] -> number of keys with digit d in the pieces of the warps before w; tot[d] = all of them; with
// kScan also dbase[d] = number of keys with a smaller digit.  Ends with a barrier.
template<int kWarps, bool kScan>
__device__ __forceinline__ void digit_prefix(uint32_t (*cnt)[kGsRadix], uint32_t* tot, uint32_t* dbase, uint32_t* wsum) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t run = 0, inc = 0;
  if(threadIdx.x < kGsRadix) {
#pragma unroll 8
    for(int w = 0; w < kWarps; ++w) { const uint32_t t = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += t; }
    tot[threadIdx.x] = run;
    if(kScan) {
      inc = run;
#pragma unroll
      for(int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(MR_FULL_MASK, inc, s); if(lane >= (unsigned)s) inc += o; }
      if(lane == 31) wsum[warp] = inc;
    }
  }
  __syncthreads();
  if(kScan) {
    if(threadIdx.x < kGsRadix) {
      uint32_t before = 0;
      for(unsigned w = 0; w < warp; ++w) before += wsum[w];
      dbase[threadIdx.x] = before + inc - run;
    }
    __syncthreads();
  }
}

// ---- a slice that does not fit shared memory: LSD passes out of global memory, ping-pong between the two
// buffer pairs, an odd number of passes so that the result lands in (kout, pout)
template<int kThreads>
__device__ void sort_slice_global(const group_sort_args& A, uint64_t seg, uint32_t n, uint32_t (*cnt)[kGsRadix],
                                  uint32_t* tot, uint32_t* gbase, uint32_t* wsum) {
  constexpr int kWarps = kThreads / 32, kChunk = kThreads * kFbRounds;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = lanemask_lt();
  int passes = (A.sr_bits + kGsDigit - 1) / kGsDigit;
  if((passes & 1) == 0) ++passes;
  const int width = (A.sr_bits + passes - 1) / passes;
  const uint32_t dmask = (1u << width) - 1;
  uint64_t *sk = A.keys_in + seg, *sp = A.pays_in + seg, *dk = A.keys_out + seg, *dp = A.pays_out + seg;
  for(int p = 0; p < passes; ++p) {
    const int shift = p * width;
    for(int i = threadIdx.x; i < kGsRadix; i += kThreads) tot[i] = 0;
    __syncthreads();
    for(uint32_t j = threadIdx.x; j < n; j += kThreads) atomicAdd(&tot[((uint32_t)sk[j] >> shift) & dmask], 1u);
    __syncthreads();
    {                                           // gbase = exclusive scan of the histogram
      uint32_t v = 0, inc = 0;
      if(threadIdx.x < kGsRadix) {
        v = tot[threadIdx.x]; inc = v;
#pragma unroll
        for(int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(MR_FULL_MASK, inc, s); if(lane >= (unsigned)s) inc += o; }
        if(lane == 31) wsum[warp] = inc;
      }
      __syncthreads();
      if(threadIdx.x < kGsRadix) {
        uint32_t before = 0;
        for(unsigned w = 0; w < warp; ++w) before += wsum[w];
        gbase[threadIdx.x] = before + inc - v;
      }
    }
    for(uint64_t c0 = 0; c0 < n; c0 += kChunk) {
      for(int i = threadIdx.x; i < kWarps * kGsRadix; i += kThreads) (&cnt[0][0])[i] = 0;
      __syncthreads();
      uint64_t key[kFbRounds], pay[kFbRounds];
      uint32_t rk[kFbRounds];
#pragma unroll
      for(int i = 0; i < kFbRounds; ++i) {
        const uint64_t pos = c0 + warp * (32 * kFbRounds) + i * 32 + lane;
        const bool valid = pos < n;
        key[i] = valid ? sk[pos] : 0; pay[i] = valid ? sp[pos] : 0;
      }
#pragma unroll
      for(int i = 0; i < kFbRounds; ++i) {
        const uint64_t pos = c0 + warp * (32 * kFbRounds) + i * 32 + lane;
        const uint32_t d = pos < n ? ((uint32_t)key[i] >> shift) & dmask : (uint32_t)kGsRadix;
        rk[i] = rank_round(d, cnt[warp], lane, lt);
      }
      __syncthreads();
      digit_prefix<kWarps, false>(cnt, tot, nullptr, nullptr);
#pragma unroll
      for(int i = 0; i < kFbRounds; ++i) {
        const uint64_t pos = c0 + warp * (32 * kFbRounds) + i * 32 + lane;
        if(pos < n) {
          const uint32_t d = ((uint32_t)key[i] >> shift) & dmask;
          const uint32_t dst = gbase[d] + cnt[warp][d] + rk[i];
          dk[dst] = key[i]; dp[dst] = pay[i];
        }
      }
      __syncthreads();
      if(threadIdx.x < kGsRadix) gbase[threadIdx.x] += tot[threadIdx.x];
      __syncthreads();
    }
    uint64_t* t = sk; sk = dk; dk = t;
    t = sp; sp = dp; dp = t;
  }
  // the sorted slice is in (sk, sp) == (keys_out, pays_out) now; its writes are visible to the whole CTA
  const uint64_t* ok = A.keys_out + seg;
  uint8_t* head = A.head + seg;
  for(uint32_t j = threadIdx.x; j < n; j += kThreads) head[j] = j == 0 || (uint32_t)ok[j] != (uint32_t)ok[j - 1];
  if(threadIdx.x == 0 && (uint32_t)ok[n - 1] == A.nseq_all) atomicAdd(A.n_invalid_groups, 1ULL);
}

template<int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 1024 ? 1 : 2) group_sort_kernel(group_sort_args A) {
  typedef gs_layout<kThreads> L;
  constexpr int kWarps = L::kWarps;
  extern __shared__ __align__(16) unsigned char gs_smem[];
  uint64_t* spay = reinterpret_cast<uint64_t*>(gs_smem + L::pay_off);
  uint32_t* arr  = reinterpret_cast<uint32_t*>(gs_smem + L::arr_off);
  uint32_t (*cnt)[kGsRadix] = reinterpret_cast<uint32_t (*)[kGsRadix]>(gs_smem + L::cnt_off);
  uint32_t* tot   = reinterpret_cast<uint32_t*>(gs_smem + L::tot_off);
  uint32_t* dbase = reinterpret_cast<uint32_t*>(gs_smem + L::dbs_off);
  uint32_t* wsum  = reinterpret_cast<uint32_t*>(gs_smem + L::wsm_off);
  uint64_t* bar   = reinterpret_cast<uint64_t*>(gs_smem + L::bar_off);

  const uint32_t r = blockIdx.x;
  const uint64_t seg = A.hit_off[A.tile_first[r]];
  const uint64_t n64 = A.hit_off[A.tile_first[r + 1]] - seg;
  if(n64 == 0) return;
  if(n64 > (uint64_t)A.cap) { sort_slice_global<kThreads>(A, seg, (uint32_t)n64, cnt, tot, dbase, wsum); return; }
  const uint32_t n = (uint32_t)n64;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = lanemask_lt();

  // payloads: one bulk copy of the slice, widened to 16-byte bounds (the hit arrays are padded by two entries)
  const uint32_t skew = (uint32_t)seg & 1u;
  if(threadIdx.x == 0) {
    const uint32_t bytes = ((n + skew + 1) & ~1u) * 8u;
    mbar_init(bar, 1);
    mbar_expect_tx(bar, bytes);
    bulk_copy_g2s(spay, A.pays_in + (seg - skew), bytes, bar);
  }
  const uint32_t ib = A.idx_bits;
  {
    const uint64_t* kin = A.keys_in + seg;
#pragma unroll 4
    for(uint32_t j = threadIdx.x; j < n; j += kThreads) arr[j] = ((uint32_t)__ldcs(kin + j) << ib) | j;
  }
  const uint32_t R = (n + kThreads - 1) / kThreads;           // rounds; warp w ranks positions [32 R w, 32 R (w + 1))
  const uint32_t piece = warp * 32 * R + lane;
  const int passes = (A.sr_bits + kGsDigit - 1) / kGsDigit;
  for(int p = 0; p < passes; ++p) {
    const int shift = (int)ib + p * kGsDigit;
    for(int i = threadIdx.x; i < kWarps * kGsRadix; i += kThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();                                           // (first pass: arr is complete, the barrier is initialised)
    uint32_t e[kGsRounds], rk[kGsRounds];
#pragma unroll
    for(int i = 0; i < kGsRounds; ++i) {
      if((uint32_t)i < R) {
        const uint32_t pos = piece + i * 32;
        const bool valid = pos < n;
        e[i] = valid ? arr[pos] : 0u;
        rk[i] = rank_round(valid ? (e[i] >> shift) & (kGsRadix - 1) : (uint32_t)kGsRadix, cnt[warp], lane, lt);
      }
    }
    __syncthreads();
    digit_prefix<kWarps, true>(cnt, tot, dbase, wsum);
#pragma unroll
    for(int i = 0; i < kGsRounds; ++i) {
      if((uint32_t)i < R && piece + i * 32 < n) {
        const uint32_t d = (e[i] >> shift) & (kGsRadix - 1);
        arr[dbase[d] + cnt[warp][d] + rk[i]] = e[i];
      }
    }
    __syncthreads();
  }
  mbar_wait(bar, 0);
  const uint64_t rhi = (uint64_t)r << 32;
  const uint32_t imask = (1u << ib) - 1;
  uint64_t* ko = A.keys_out + seg;
  uint64_t* po = A.pays_out + seg;
  uint8_t* head = A.head + seg;
#pragma unroll 4
  for(uint32_t j = threadIdx.x; j < n; j += kThreads) {
    const uint32_t x = arr[j], sr = x >> ib;
    ko[j] = rhi | sr;
    po[j] = spay[skew + (x & imask)];
    head[j] = j == 0 || (arr[j - 1] >> ib) != sr;
  }
  if(threadIdx.x == 0 && (arr[n - 1] >> ib) == A.nseq_all) atomicAdd(A.n_invalid_groups, 1ULL);
}

// hits held by reads whose slice does not fit the shared-memory sort
__global__ void __launch_bounds__(256) read_hits_stats_kernel(const uint64_t* __restrict__ hit_off, const uint32_t* __restrict__ tile_first,
                                                               uint32_t nreads, uint32_t cap, unsigned long long* __restrict__ big_hits) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n = 0;
  if(r < nreads) { n = hit_off[tile_first[r + 1]] - hit_off[tile_first[r]]; if(n <= cap) n = 0; }
  for(int s = 16; s > 0; s >>= 1) n += __shfl_down_sync(MR_FULL_MASK, n, s);
  if((threadIdx.x & 31) == 0 && n) atomicAdd(big_hits, n);
}

} // namespace

int group_sort_threads() {
  static const int v = [] { const char* e = getenv("MR_GSORT_THREADS"); const int x = e ? atoi(e) : 0; return x == 512 ? 512 : 1024; }();
  return v;
}

// MR_GSORT_CAP lowers the capacity (a power of two, at least 32): the tests send the fixtures' reads through the
// global-memory route that way
uint32_t group_sort_capacity(int sr_bits) {
  static const int env_bits = [] { const char* e = getenv("MR_GSORT_CAP"); int x = e ? atoi(e) : 0, b = 0; while((2 << b) <= x) ++b; return x >= 32 ? b : 31; }();
  const int ib = std::min(std::min(32 - sr_bits, group_sort_threads() == 1024 ? 14 : 13), env_bits);
  return ib < 5 ? 0u : 1u << ib;
}

int launch_read_hits_stats(mr_context* ctx, const uint64_t* hit_off, const uint32_t* tile_first, uint32_t nreads, uint32_t cap,
                           unsigned long long* big_hits) {
  if(!nreads) return MR_OK;
  read_hits_stats_kernel<<<div_up(nreads, 256), 256, 0, ctx->stream>>>(hit_off, tile_first, nreads, cap, big_hits);
  MR_LAUNCHED(ctx);
  return MR_OK;
}

int launch_group_sort(mr_context* ctx, group_sort_args A, uint32_t nreads) {
  if(!nreads) return MR_OK;
  A.cap = group_sort_capacity(A.sr_bits);
  if(!A.cap) return ctx->fail(MR_ELIMIT, "group sort: super-read index too wide for the per-read sort");
  A.idx_bits = 0;
  while((1u << A.idx_bits) < A.cap) ++A.idx_bits;
  if(group_sort_threads() == 1024) {
    MR_CUDA(ctx, cudaFuncSetAttribute(group_sort_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs_layout<1024>::bytes));
    group_sort_kernel<1024><<<nreads, 1024, gs_layout<1024>::bytes, ctx->stream>>>(A);
  } else {
    MR_CUDA(ctx, cudaFuncSetAttribute(group_sort_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs_layout<512>::bytes));
    group_sort_kernel<512><<<nreads, 512, gs_layout<512>::bytes, ctx->stream>>>(A);
  }
  MR_LAUNCHED(ctx);
  return MR_OK;
}
