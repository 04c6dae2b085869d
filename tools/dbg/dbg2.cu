#include "../../mindthegap_b200/csrc/common.cuh"
#include <vector>
using namespace mtg;
__global__ void k1(const u128* keys, int n, int k, uint64_t* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u128 x = keys[i];
    uint64_t a = rc_word((uint64_t)x), b = rc_word((uint64_t)(x >> 64));
    u128 r = ((u128)a << 64) | (u128)b;
    int s = 2 * (64 - k);
    u128 sh = r >> s;
    uint64_t manual_lo = (b >> s) | (a << (64 - s));
    uint64_t br = __brevll((uint64_t)x);
    out[i * 8 + 0] = a; out[i * 8 + 1] = b; out[i * 8 + 2] = (uint64_t)sh; out[i * 8 + 3] = (uint64_t)(sh >> 64); out[i * 8 + 4] = manual_lo; out[i*8+5] = br;
    u128 r2 = revcomp(x, k);
    out[i*8+6] = (uint64_t)r2; out[i*8+7] = (uint64_t)(r2>>64);
}
int main() {
    int n = 2, k = 61;
    std::vector<u128> keys(n);
    keys[0] = ((u128)0x0277b20aec4233f8ULL << 64) | 0xf277723c109dd69cULL;
    keys[1] = ((u128)0x03494d6880418a99ULL << 64) | 0xe555d4ed5e64cfd3ULL;
    u128* d; uint64_t *dout;
    cudaMalloc(&d, n * 16); cudaMalloc(&dout, n * 64);
    cudaMemcpy(d, keys.data(), n * 16, cudaMemcpyHostToDevice);
    k1<<<1, 32>>>(d, n, k, dout);
    std::vector<uint64_t> out(n * 8);
    cudaMemcpy(out.data(), dout, n * 64, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; i++) {
        u128 x = keys[i];
        uint64_t a = rc_word((uint64_t)x), b = rc_word((uint64_t)(x >> 64));
        u128 r = revcomp(x, k);
        printf("dev a %016lx b %016lx sh %016lx%016lx manual_lo %016lx brev %016lx rc %016lx%016lx\n", out[i*8], out[i*8+1], out[i*8+3], out[i*8+2], out[i*8+4], out[i*8+5], out[i*8+7], out[i*8+6]);
        printf("hst a %016lx b %016lx rc %016lx%016lx\n", a, b, (uint64_t)(r>>64), (uint64_t)r);
    }
}
