// primitives.cuh -- device-wide exclusive scan and a stable LSD radix sort, both
// hand-written (no CUB/Thrust on the path).  Deterministic: identical inputs give
// identical outputs, which the bit-exact parity tests rely on.
#pragma once

#include "common.cuh"

namespace arcte {

// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 entries (out[n] = total).
// `scratch` is grown as needed.  Counts launched kernels into *launches.
int exclusive_scan_i32(const int32_t *in, int64_t *out, int64_t n, DevBuf &scratch,
                       cudaStream_t stream, int64_t *launches);
int exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, DevBuf &scratch,
                       cudaStream_t stream, int64_t *launches);

// Stable sort of n (key, value) pairs by the low `bits` bits of the uint32 key.
// Ping-pongs between (k0, v0) and (k1, v1); *result_in_second tells where the
// sorted data ended up.  value_bytes is 4 or 8.
int radix_sort_pairs(uint32_t *k0, void *v0, uint32_t *k1, void *v1, int64_t n, int bits,
                     int value_bytes, DevBuf &scratch_hist, DevBuf &scratch_scan,
                     DevBuf &scratch_scan2, cudaStream_t stream, bool *result_in_second,
                     int64_t *launches);

}  // namespace arcte
