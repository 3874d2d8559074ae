// Probe: is max(-lg2.approx(u) * ln2, 2^-24) bit-identical to torch's guarded form
//   u >= 1 - 2^-25 ? 2^-24 : -__logf(u)
// for every 32-bit Philox word?  (exhaustive over the top 2^20 words, where the two can differ, plus a stride sample)
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float u_of(uint32_t x) { return fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f); }
__device__ __forceinline__ float lg2a(float u) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u)); return r; }
__global__ void probe(unsigned long long* bad, uint32_t base, uint32_t stride) {
    const uint32_t x = base + (blockIdx.x * blockDim.x + threadIdx.x) * stride;
    const float u = u_of(x);
    const float l = __fmul_rn(lg2a(u), 0.693147182464599609375f);
    const float ref = (u >= 1.0f - 1.1920928955078125e-07f / 2.0f) ? (1.1920928955078125e-07f / 2.0f) : -l;
    const float alt = fmaxf(-l, 1.1920928955078125e-07f / 2.0f);
    if (__float_as_uint(ref) != __float_as_uint(alt)) atomicAdd(bad, 1ull);
}
int main() {
    unsigned long long *bad, h = 0;
    cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    probe<<<(1 << 20) / 256, 256>>>(bad, 0xFFF00000u, 1u);           // top 2^20 words, exhaustive
    probe<<<(1 << 24) / 256, 256>>>(bad, 0u, 256u);                  // every 256th word of the whole range
    cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
    printf("mismatches %llu\n", h);
    return 0;
}
