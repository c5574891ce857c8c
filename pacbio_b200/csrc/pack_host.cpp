// Host side of the read transport: ASCII bases -> 2 bits per base + one mask bit per base (the layout of
// compact_dna.hpp:102-136, A0 C1 G2 T3; everything outside ACGTacgt sets the mask bit and packs as 0, the k-mer
// breaks of jf_aligner.hpp:41-52).  This runs on every base that reaches a GPU, so it has to keep up with the
// GPUs of the box: a byte-at-a-time table lookup packs 0.46 GB/s per core -- 1.3 core-seconds for the 0.6 GB
// of reads one B200 aligns in 75 ms, the whole reason the 8-GPU end-to-end rate stopped scaling -- the AVX2
// form below 32 bases per dozen instructions.  Chosen at run time (the library is not built with -mavx2).
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <immintrin.h>

#include "../../include/mega_reads_b200.h"

namespace {

const struct lut_t {
  uint8_t v[256];
  lut_t() { for(int i = 0; i < 256; ++i) v[i] = 4; v['a'] = v['A'] = 0; v['c'] = v['C'] = 1; v['g'] = v['G'] = 2; v['t'] = v['T'] = 3; }
} lut;

// 64 characters -> two code words and one mask word, one character at a time
inline void pack64_scalar(const unsigned char* s, uint64_t& c0, uint64_t& c1, uint64_t& m) {
  c0 = c1 = m = 0;
  for(int j = 0; j < 32; ++j) { const uint64_t c = lut.v[s[j]]; c0 |= (c & 3) << (2 * j); m |= (c >> 2) << j; }
  for(int j = 0; j < 32; ++j) { const uint64_t c = lut.v[s[32 + j]]; c1 |= (c & 3) << (2 * j); m |= (c >> 2) << (32 + j); }
}

// 32 characters at once.  (c >> 1) & 3 sends A C G T (either case) to 0 1 3 2; x ^ (x >> 1) turns that into 0 1 2 3.
// Pairs of 2-bit values are merged by multiply-adds (1 and 4, then 1 and 16), which leaves 4 bases in the low byte
// of every 32-bit lane; a byte shuffle gathers the eight bytes.
__attribute__((target("avx2")))
inline void pack32_avx2(const unsigned char* s, uint64_t& code, uint32_t& invalid) {
  const __m256i v  = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
  const __m256i up = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));
  const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'))),
                                     _mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(up, _mm256_set1_epi8('T'))));
  invalid = ~(uint32_t)_mm256_movemask_epi8(ok);
  const __m256i t = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));
  __m256i c = _mm256_xor_si256(t, _mm256_and_si256(_mm256_srli_epi16(t, 1), _mm256_set1_epi8(1)));
  c = _mm256_and_si256(c, ok);
  const __m256i p16 = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0401));
  const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi32(0x00100001));
  const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  const __m256i g = _mm256_shuffle_epi8(p32, pick);
  code = (uint64_t)(uint32_t)_mm256_extract_epi32(g, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(g, 4) << 32);
}

__attribute__((target("avx2")))
void pack_groups_avx2(const unsigned char* s, uint64_t groups, uint64_t* codes, uint64_t* nmask) {
  for(uint64_t w = 0; w < groups; ++w, s += 64) {
    uint32_t i0, i1;
    pack32_avx2(s, codes[2 * w], i0);
    pack32_avx2(s + 32, codes[2 * w + 1], i1);
    nmask[w] = (uint64_t)i0 | ((uint64_t)i1 << 32);
  }
}

void pack_groups_scalar(const unsigned char* s, uint64_t groups, uint64_t* codes, uint64_t* nmask) {
  for(uint64_t w = 0; w < groups; ++w, s += 64) pack64_scalar(s, codes[2 * w], codes[2 * w + 1], nmask[w]);
}

// MR_PACK_SCALAR=1: the portable form (A/B, tests)
const bool g_avx2 = !(getenv("MR_PACK_SCALAR") && atoi(getenv("MR_PACK_SCALAR")) != 0) && __builtin_cpu_supports("avx2");

} // namespace

extern "C" {

uint64_t mr_packed_code_words(uint64_t nbases) { return (nbases + 31) / 32 + 6; }
uint64_t mr_packed_mask_words(uint64_t nbases) { return (nbases + 63) / 64 + 6; }

// mr_pack_reads_range fills mask words [first_word, first_word + n_words) and the code words that go with them,
// so that several host threads can pack disjoint word ranges of one batch.
int mr_pack_reads_range(const char* bases, uint64_t nbases, uint64_t first_word, uint64_t n_words, uint64_t* codes, uint64_t* nmask) {
  if((!bases && nbases) || !codes || !nmask) return MR_EINVAL;
  const uint64_t cwords = mr_packed_code_words(nbases), mwords = mr_packed_mask_words(nbases);
  const uint64_t end_word = std::min(mwords, first_word + n_words);
  const uint64_t full = nbases / 64;                       // words whose 64 characters all exist
  uint64_t w = first_word;
  if(w < std::min(end_word, full)) {
    const uint64_t n = std::min(end_word, full) - w;
    const unsigned char* s = reinterpret_cast<const unsigned char*>(bases) + w * 64;
    if(g_avx2) pack_groups_avx2(s, n, codes + 2 * w, nmask + w);
    else       pack_groups_scalar(s, n, codes + 2 * w, nmask + w);
    w += n;
  }
  for(; w < end_word; ++w) {                               // the partial group and the padding words
    uint64_t c0 = 0, c1 = 0, m = 0;
    const uint64_t g = w * 64;
    for(uint64_t j = 0; j < 64 && g + j < nbases; ++j) {
      const uint64_t c = lut.v[(unsigned char)bases[g + j]];
      if(j < 32) c0 |= (c & 3) << (2 * j); else c1 |= (c & 3) << (2 * (j - 32));
      m |= (c >> 2) << j;
    }
    if(2 * w < cwords) codes[2 * w] = c0;
    if(2 * w + 1 < cwords) codes[2 * w + 1] = c1;
    nmask[w] = m;
  }
  return MR_OK;
}

int mr_pack_reads(const char* bases, uint64_t nbases, uint64_t* codes, uint64_t* nmask) {
  return mr_pack_reads_range(bases, nbases, 0, mr_packed_mask_words(nbases), codes, nmask);
}

}
