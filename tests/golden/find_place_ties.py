"""Find (group key, d1, d2) with equal rendezvous scores mix32(key ^ d·φ)|1 for the placement
tie-break tests (tests/test_place_spec.py PLACE_TIES): invert the murmur3 finalizer for y and y^1
and look for a difference that equals d1·φ ^ d2·φ for small d1 < d2."""
import numpy as np

M = 0xFFFFFFFF
PHI = 0x9E3779B1


def mix(x):
    x = x.astype(np.uint64)
    x ^= x >> 16; x = (x * 0x85EBCA6B) & M; x ^= x >> 13; x = (x * 0xC2B2AE35) & M; x ^= x >> 16
    return x


def inv(y):
    y = y.astype(np.uint64)
    y ^= y >> 16; y = (y * 0x7ED1B41D) & M; y ^= (y >> 13) ^ (y >> 26); y = (y * 0xA5CB9243) & M; y ^= y >> 16
    return y


if __name__ == "__main__":
    D = 64
    dphi = [(d * PHI) & M for d in range(D)]
    pairs = {dphi[a] ^ dphi[b]: (a, b) for a in range(D) for b in range(a + 1, D)}
    keys = np.array(sorted(pairs), dtype=np.uint64)
    found = []
    for blk in range(64):
        y = (np.arange(blk << 22, (blk + 1) << 22, dtype=np.uint64) * 2 + 1) & M
        xa, xb = inv(y), inv(y ^ 1)
        delta = xa ^ xb
        idx = np.minimum(np.searchsorted(keys, delta), len(keys) - 1)
        for i in np.flatnonzero(keys[idx] == delta):
            a, b = pairs[int(delta[i])]
            key = int(xa[i]) ^ dphi[a]
            score = int(mix(np.array([key ^ dphi[a]], dtype=np.uint64))[0]) | 1
            assert score == int(mix(np.array([key ^ dphi[b]], dtype=np.uint64))[0]) | 1
            found.append((hex(key), a, b, hex(score)))
        if len(found) >= 5:
            break
    print(found)
