// see dsk_export.cuh. Off the timed path: the ordering uses CUB's device radix sort (library code, like cuBLAS for a GEMM).
#include "dsk_export.cuh"

#include <algorithm>
#include <queue>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

namespace mtg {
namespace {

template <class K>
__global__ void __launch_bounds__(256) minimizer_kernel(const K* __restrict__ keys, uint64_t n, int k, int m, uint32_t* __restrict__ mm,
                                                        unsigned int* __restrict__ hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = gatb_minimizer<K>(keys[i], k, m);
        mm[i] = v;
        atomicAdd(hist + v, 1u);
    }
}
__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ idx, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) idx[i] = (uint32_t)i;
}
// sort key of pass `which` for the current order: 0 low word, 1 high word, 2 partition of the k-mer's minimizer
template <class K>
__global__ void __launch_bounds__(256) field_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ mm, const uint16_t* __restrict__ repart,
                                                    const uint32_t* __restrict__ idx, uint64_t n, int which, uint64_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t j = idx[i];
        out[i] = which == 0 ? lo64(keys[j]) : which == 1 ? hi64(keys[j]) : (uint64_t)repart[mm[j]];
    }
}
template <class K>
__global__ void __launch_bounds__(256) gather_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ abund, const uint32_t* __restrict__ idx,
                                                     uint64_t n, uint64_t* __restrict__ lo, uint64_t* __restrict__ hi, uint32_t* __restrict__ ab) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t j = idx[i];
        lo[i] = lo64(keys[j]);
        hi[i] = hi64(keys[j]);
        ab[i] = abund ? abund[j] : 0u;
    }
}
inline int grid_of(uint64_t n) { return (int)std::min<uint64_t>(std::max<uint64_t>((n + 255) / 256, 1), 148 * 16); }

}  // namespace

template <class K>
void dsk_partition_export(const K* d_keys, const uint32_t* d_abund, uint64_t n, int k, int m, uint32_t nparts, cudaStream_t stream,
                          uint16_t* repart, uint64_t* part_offsets, uint64_t* lo, uint64_t* hi, uint32_t* abundance) {
    if (m < 3 || m > 12 || m >= k) throw Error(-1, "dsk export: minimizer size must be in [3,12] and below k");
    if (nparts < 1 || nparts > 65535) throw Error(-1, "dsk export: 1..65535 partitions");
    if (n >= (1ull << 32)) throw Error(-1, "dsk export: more than 2^32 solid k-mers per call (export the shares of several GPUs one by one)");
    const uint64_t nmin = 1ull << (2 * m);
    // 1. minimizer of every solid k-mer + solid k-mers per minimizer
    DevBuf<uint32_t> mm(std::max<uint64_t>(n, 1));
    DevBuf<unsigned int> hist(nmin);
    hist.zero(stream);
    if (n) minimizer_kernel<K><<<grid_of(n), 256, 0, stream>>>(d_keys, n, k, m, mm.p, hist.p);
    MTG_CUDA(cudaGetLastError());
    std::vector<unsigned int> h(nmin);
    MTG_CUDA(cudaMemcpyAsync(h.data(), hist.p, nmin * 4, cudaMemcpyDeviceToHost, stream));
    MTG_CUDA(cudaStreamSynchronize(stream));
    // 2. Repartitor::computeDistrib: bins by decreasing size, each into the partition with the least load so far
    std::vector<std::pair<uint64_t, uint32_t>> bins(nmin);
    for (uint64_t i = 0; i < nmin; i++) bins[i] = std::make_pair((uint64_t)h[i], (uint32_t)i);
    std::stable_sort(bins.begin(), bins.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first > b.first; });
    typedef std::pair<uint64_t, uint32_t> Load;   // (load, partition): smallest load first, then smallest partition id
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pq;
    for (uint32_t p = 0; p < nparts; p++) pq.push(Load(0, p));
    std::vector<uint64_t> psize(nparts, 0);
    for (uint64_t i = 0; i < nmin; i++) {
        Load s = pq.top(); pq.pop();
        repart[bins[i].second] = (uint16_t)s.second;
        s.first += bins[i].first;
        psize[s.second] += bins[i].first;
        pq.push(s);
    }
    part_offsets[0] = 0;
    for (uint32_t p = 0; p < nparts; p++) part_offsets[p + 1] = part_offsets[p] + psize[p];
    if (!n) return;
    // 3. order by (partition, k-mer): stable LSD passes over (low word, high word, partition)
    DevBuf<uint16_t> d_repart(nmin);
    MTG_CUDA(cudaMemcpyAsync(d_repart.p, repart, nmin * 2, cudaMemcpyHostToDevice, stream));
    DevBuf<uint32_t> idx_a(n), idx_b(n);
    DevBuf<uint64_t> f_a(n), f_b(n);
    iota_kernel<<<grid_of(n), 256, 0, stream>>>(idx_a.p, n);
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, f_a.p, f_b.p, idx_a.p, idx_b.p, (int)n, 0, 64, stream);
    DevBuf<uint8_t> temp(temp_bytes + 16);
    uint32_t *cur = idx_a.p, *nxt = idx_b.p;
    const int passes[3] = {0, 1, 2};
    for (int pi = 0; pi < 3; pi++) {
        const int which = passes[pi];
        if (which == 1 && sizeof(K) == 8) continue;
        field_kernel<K><<<grid_of(n), 256, 0, stream>>>(d_keys, mm.p, d_repart.p, cur, n, which, f_a.p);
        MTG_CUDA(cudaGetLastError());
        size_t tb = temp_bytes;
        const int end_bit = which == 2 ? 16 : 64;
        MTG_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, f_a.p, f_b.p, cur, nxt, (int)n, 0, end_bit, stream));
        std::swap(cur, nxt);
    }
    // 4. gather in that order, copy to the host arrays
    DevBuf<uint64_t> o_lo(n), o_hi(n);
    DevBuf<uint32_t> o_ab(n);
    gather_kernel<K><<<grid_of(n), 256, 0, stream>>>(d_keys, d_abund, cur, n, o_lo.p, o_hi.p, o_ab.p);
    MTG_CUDA(cudaGetLastError());
    MTG_CUDA(cudaMemcpyAsync(lo, o_lo.p, n * 8, cudaMemcpyDeviceToHost, stream));
    if (hi) MTG_CUDA(cudaMemcpyAsync(hi, o_hi.p, n * 8, cudaMemcpyDeviceToHost, stream));
    if (abundance) MTG_CUDA(cudaMemcpyAsync(abundance, o_ab.p, n * 4, cudaMemcpyDeviceToHost, stream));
    MTG_CUDA(cudaStreamSynchronize(stream));
}

template void dsk_partition_export<uint64_t>(const uint64_t*, const uint32_t*, uint64_t, int, int, uint32_t, cudaStream_t, uint16_t*, uint64_t*, uint64_t*,
                                             uint64_t*, uint32_t*);
template void dsk_partition_export<u128>(const u128*, const uint32_t*, uint64_t, int, int, uint32_t, cudaStream_t, uint16_t*, uint64_t*, uint64_t*, uint64_t*,
                                         uint32_t*);

}  // namespace mtg
