// Host-side helpers of the Python host layer's Fiat-Shamir generator (ckb_zkp_b200/fs_rng.py restates
// marlin/src/fs_rng.rs): the Keccak-f[1600] permutation under STROBE-128 / Merlin and ChaCha20 keystream blocks
// in rand_chacha's layout.  Plain C++ on the host -- a Rust host gets both from the merlin and rand_chacha crates
// and never calls these.  No device code, no GPU needed.
#include <cstddef>
#include <cstdint>

#include "../../include/zkb.h"

namespace {

inline uint64_t rol64(uint64_t x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }
inline uint32_t rol32(uint32_t x, unsigned n) { return (x << n) | (x >> (32 - n)); }

const uint64_t kRoundConstants[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
// rho offsets and pi destinations in the order of the classic in-place lane walk starting at lane 1
const unsigned kRho[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
const unsigned kPi[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

}  // namespace

extern "C" {

void zkb_host_keccak_f1600(uint64_t st[25]) {
  for (int round = 0; round < 24; round++) {
    uint64_t bc[5];
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) {
      uint64_t t = bc[(i + 4) % 5] ^ rol64(bc[(i + 1) % 5], 1);
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    uint64_t t = st[1];
    for (int i = 0; i < 24; i++) {
      unsigned j = kPi[i];
      uint64_t b = st[j];
      st[j] = rol64(t, kRho[i]);
      t = b;
    }
    for (int j = 0; j < 25; j += 5) {
      for (int i = 0; i < 5; i++) bc[i] = st[j + i];
      for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    st[0] ^= kRoundConstants[round];
  }
}

// n_blocks consecutive 64-byte ChaCha20 blocks (20 rounds) as little-endian u32 words: key = 8 words, words 12-13 =
// 64-bit block counter starting at `counter`, words 14-15 = 0 (rand_chacha 0.2 ChaChaRng::from_seed, stream 0)
void zkb_host_chacha20_blocks(const uint8_t key[32], uint64_t counter, uint32_t* out_words, size_t n_blocks) {
  uint32_t k[8];
  for (int i = 0; i < 8; i++)
    k[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
  for (size_t b = 0; b < n_blocks; b++, counter++) {
    uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7],
                         (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t s[16];
    for (int i = 0; i < 16; i++) s[i] = init[i];
#define ZKB_QR(a, b, c, d)                                    \
  s[a] += s[b]; s[d] = rol32(s[d] ^ s[a], 16); s[c] += s[d]; s[b] = rol32(s[b] ^ s[c], 12); \
  s[a] += s[b]; s[d] = rol32(s[d] ^ s[a], 8);  s[c] += s[d]; s[b] = rol32(s[b] ^ s[c], 7);
    for (int r = 0; r < 10; r++) {
      ZKB_QR(0, 4, 8, 12) ZKB_QR(1, 5, 9, 13) ZKB_QR(2, 6, 10, 14) ZKB_QR(3, 7, 11, 15)
      ZKB_QR(0, 5, 10, 15) ZKB_QR(1, 6, 11, 12) ZKB_QR(2, 7, 8, 13) ZKB_QR(3, 4, 9, 14)
    }
#undef ZKB_QR
    for (int i = 0; i < 16; i++) out_words[16 * b + i] = s[i] + init[i];
  }
}

}  // extern "C"
