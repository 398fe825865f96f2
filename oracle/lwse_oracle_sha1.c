/*
 * lwse_oracle_sha1.c — CPU SHA-1 for the group / subgroup keys.
 *
 * TEST INFRASTRUCTURE ONLY (see lwse_oracle.c).  Restates pkg/utils/utils.go:39-43
 * Sha1Hash = hex(crypto/sha1 of the string) — Go's crypto/sha1 is FIPS 180-4 SHA-1
 * (stdlib, not under /root/reference).  Pinned by the three known answers of
 * pkg/webhooks/pod_webhook_test.go:29-53 (tests/test_oracle_golden.py) and
 * cross-checked against Python's hashlib in the same test.
 */
#include <stdint.h>
#include <string.h>

#include "../include/lwse.h"

#define LWSO_API __attribute__((visibility("default")))

static uint32_t rol(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

static void block(uint32_t h[5], const uint8_t* p) {
  uint32_t w[80];
  for (int t = 0; t < 16; t++)
    w[t] = ((uint32_t)p[4 * t] << 24) | ((uint32_t)p[4 * t + 1] << 16) | ((uint32_t)p[4 * t + 2] << 8) | p[4 * t + 3];
  for (int t = 16; t < 80; t++) w[t] = rol(w[t - 3] ^ w[t - 8] ^ w[t - 14] ^ w[t - 16], 1);
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
  for (int t = 0; t < 80; t++) {
    uint32_t f, k;
    if (t < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
    else if (t < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
    else if (t < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
    else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
    uint32_t tmp = rol(a, 5) + f + e + k + w[t];
    e = d; d = c; c = rol(b, 30); b = a; a = tmp;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

LWSO_API int lwso_sha1(const uint8_t* bytes, const uint32_t* offsets, uint32_t n, uint8_t* digests) {
  for (uint32_t i = 0; i < n; i++) {
    const uint8_t* msg = bytes + offsets[i];
    uint32_t len = offsets[i + 1] - offsets[i];
    uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    uint32_t full = len / 64;
    for (uint32_t k = 0; k < full; k++) block(h, msg + 64 * k);
    uint8_t tail[128];
    uint32_t rem = len - 64 * full;
    memset(tail, 0, sizeof(tail));
    memcpy(tail, msg + 64 * full, rem);
    tail[rem] = 0x80;
    uint32_t tl = rem + 9 <= 64 ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int b = 0; b < 8; b++) tail[tl - 1 - b] = (uint8_t)(bits >> (8 * b));
    block(h, tail);
    if (tl == 128) block(h, tail + 64);
    for (int j = 0; j < 5; j++) {
      digests[i * 20 + 4 * j] = (uint8_t)(h[j] >> 24);
      digests[i * 20 + 4 * j + 1] = (uint8_t)(h[j] >> 16);
      digests[i * 20 + 4 * j + 2] = (uint8_t)(h[j] >> 8);
      digests[i * 20 + 4 * j + 3] = (uint8_t)h[j];
    }
  }
  return LWSE_OK;
}
