// SHA-1 group / subgroup keys (sm_100a): one thread per key string.
//
// pkg/webhooks/pod_webhook.go:180-182 genGroupUniqueKey = Sha1Hash("<ns>/<podName>")
// and :130,:151 Sha1Hash("<leaderName>/<subGroupIndex>"); pkg/utils/utils.go:39-43
// Sha1Hash = hex(crypto/sha1).  The engine emits the 20 raw digest bytes; the
// caller hex-encodes (the label value is the 40-character hex string).
// FIPS 180-4 SHA-1, big-endian message schedule kept in a 16-word ring.
#include "lwse_device.cuh"

namespace lwse {

__device__ __forceinline__ uint32_t rotl(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

__device__ __forceinline__ void sha1_block(uint32_t* h, uint32_t* w) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
#pragma unroll
  for (int t = 0; t < 80; t++) {
    if (t >= 16) {
      const uint32_t x = w[(t - 3) & 15] ^ w[(t - 8) & 15] ^ w[(t - 14) & 15] ^ w[t & 15];
      w[t & 15] = rotl(x, 1);
    }
    uint32_t f, k;
    if (t < 20) {
      f = (b & c) | (~b & d);
      k = 0x5A827999u;
    } else if (t < 40) {
      f = b ^ c ^ d;
      k = 0x6ED9EBA1u;
    } else if (t < 60) {
      f = (b & c) | (b & d) | (c & d);
      k = 0x8F1BBCDCu;
    } else {
      f = b ^ c ^ d;
      k = 0xCA62C1D6u;
    }
    const uint32_t tmp = rotl(a, 5) + f + e + k + w[t & 15];
    e = d;
    d = c;
    c = rotl(b, 30);
    b = a;
    a = tmp;
  }
  h[0] += a;
  h[1] += b;
  h[2] += c;
  h[3] += d;
  h[4] += e;
}

__global__ void __launch_bounds__(128) sha1_kernel(const uint8_t* __restrict__ bytes,
                                                   const uint32_t* __restrict__ offsets, uint32_t n,
                                                   uint8_t* __restrict__ digests) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t begin = __ldg(offsets + i), end = __ldg(offsets + i + 1);
    const uint32_t len = end - begin;
    const uint8_t* msg = bytes + begin;
    uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    uint32_t w[16];
    // message bytes, then 0x80, zeros, and the 64-bit big-endian bit length
    const uint32_t total = ((len + 8u) / 64u + 1u) * 64u;
    for (uint32_t pos = 0; pos < total; pos += 64u) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const uint32_t p = pos + (uint32_t)j * 4u + (uint32_t)b;
          uint32_t byte;
          if (p < len)
            byte = __ldg(msg + p);
          else if (p == len)
            byte = 0x80u;
          else
            byte = 0u;
          word = (word << 8) | byte;
        }
        w[j] = word;
      }
      if (pos + 64u == total) {
        w[14] = len >> 29;  // bit length, high word
        w[15] = len << 3;
      }
      sha1_block(h, w);
    }
    uint8_t* out = digests + (size_t)i * 20u;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      out[j * 4 + 0] = (uint8_t)(h[j] >> 24);
      out[j * 4 + 1] = (uint8_t)(h[j] >> 16);
      out[j * 4 + 2] = (uint8_t)(h[j] >> 8);
      out[j * 4 + 3] = (uint8_t)h[j];
    }
  }
}

int launch_sha1(const uint8_t* d_bytes, const uint32_t* d_offsets, uint32_t n, uint8_t* d_digests,
                int sm_count, cudaStream_t s, int* cuda_err) {
  *cuda_err = 0;
  if (n == 0) return 0;
  const uint32_t want = (n + 127u) / 128u;
  const uint32_t cap = (uint32_t)sm_count * 16u;
  sha1_kernel<<<want < cap ? want : cap, 128, 0, s>>>(d_bytes, d_offsets, n, d_digests);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *cuda_err = (int)e;
    return -1;
  }
  return 1;
}

}  // namespace lwse
