// tdt_sort.cuh -- stable LSD radix sort of (key, int32 value) pairs on the low `bits` key bits.
//
// Round-1 first version: the device-wide sort is CUB's onesweep (library code, like cuBLAS for a
// plain GEMM); everything around it is ours.  The interface is what the hand-written sm_100a
// onesweep replaces it behind.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "tdt_common.cuh"

namespace tdt {

// upper bound of the temporary storage sort_pairs needs for n pairs (no device query: callable
// without a GPU).  CUB onesweep: per-pass digit histograms + one 256-entry look-back row per tile.
static inline size_t sort_temp_bytes(int64_t n) {
    return (size_t)(n / 2) + ((size_t)4 << 20);
}

template <typename K>
static int sort_pairs(K *keys_a, K *keys_b, int32_t *vals_a, int32_t *vals_b, int64_t n, int bits, void *temp,
                      size_t temp_bytes, cudaStream_t st, int *which) {
    cub::DoubleBuffer<K> dk(keys_a, keys_b);
    cub::DoubleBuffer<int32_t> dv(vals_a, vals_b);
    if (bits <= 0) {  // nothing to order: already "sorted", stable
        *which = 0;
        return TDT_OK;
    }
    size_t need = 0;
    TDT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, n, 0, bits, st));
    if (need > temp_bytes) return fail(TDT_E_WORKSPACE, "sort needs %zu bytes of temporary storage, %zu reserved", need, temp_bytes);
    TDT_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, dk, dv, n, 0, bits, st));
    *which = dk.Current() == keys_a ? 0 : 1;
    if ((dv.Current() == vals_a ? 0 : 1) != *which) return fail(TDT_E_CUDA, "radix sort left keys and values in different buffers");
    return TDT_OK;
}

}  // namespace tdt
