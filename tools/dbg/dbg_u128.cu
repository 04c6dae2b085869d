#include "../../mindthegap_b200/csrc/graph.cuh"
#include <vector>
using namespace mtg;
__global__ void k1(const u128* keys, int n, int k, uint64_t tai, uint64_t seed0, const uint64_t* rnd, uint64_t* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u128 item = keys[i];
    u128 hashpart = (item >> 2) & kmask<u128>(k - 2);
    u128 rev = revcomp(hashpart, k - 2);
    out[i * 8 + 0] = (uint64_t)hashpart; out[i * 8 + 1] = (uint64_t)(hashpart >> 64);
    out[i * 8 + 2] = (uint64_t)rev; out[i * 8 + 3] = (uint64_t)(rev >> 64);
    out[i * 8 + 4] = rev < hashpart;
    if (rev < hashpart) hashpart = rev;
    out[i * 8 + 5] = gatb_hash1(hashpart, seed0);
    uint64_t h[8];
    bloom_neighbor_positions<u128>(k, tai, 4, seed0, rnd, item, h);
    out[i * 8 + 6] = h[0]; out[i * 8 + 7] = h[3];
}
int main() {
    int n = 8, k = 63;
    std::vector<u128> keys(n);
    uint64_t x = 88172645463325252ULL;
    for (int i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; uint64_t a = x; x ^= x << 13; x ^= x >> 7; x ^= x << 17; keys[i] = (((u128)(a >> 2)) << 64) | x; }
    u128* d; uint64_t *dout, *drnd;
    cudaMalloc(&d, n * 16); cudaMalloc(&dout, n * 64); cudaMalloc(&drnd, 2048);
    cudaMemcpy(d, keys.data(), n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(drnd, MTG_RANDOM_VALUES, 2048, cudaMemcpyHostToDevice);
    uint64_t seed0 = 0xAAAAAAAA55555555ULL * 0xB5B5B5B54B4B4B4BULL, tai = 49603;
    k1<<<1, 32>>>(d, n, k, tai, seed0, drnd, dout);
    std::vector<uint64_t> out(n * 8);
    cudaMemcpy(out.data(), dout, n * 64, cudaMemcpyDeviceToHost);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < n; i++) {
        u128 item = keys[i];
        u128 hashpart = (item >> 2) & kmask<u128>(k - 2);
        u128 rev = revcomp(hashpart, k - 2);
        int lt = rev < hashpart;
        u128 hp = lt ? rev : hashpart;
        uint64_t hh = gatb_hash1(hp, seed0);
        printf("%d dev hp %016lx%016lx rev %016lx%016lx lt %lu hash %016lx h0 %lu h3 %lu\n", i, out[i*8+1], out[i*8], out[i*8+3], out[i*8+2], out[i*8+4], out[i*8+5], out[i*8+6], out[i*8+7]);
        printf("%d hst hp %016lx%016lx rev %016lx%016lx lt %d hash %016lx\n", i, (uint64_t)(hashpart>>64), (uint64_t)hashpart, (uint64_t)(rev>>64), (uint64_t)rev, lt, hh);
    }
    return 0;
}
