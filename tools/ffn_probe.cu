// Developer probe: times the fused FFN kernel alone and prints the in-kernel timeline of CTA 0 (-DMESM_TC_TIMING).
#include "../mesm_b200/csrc/kernels.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace mesm { thread_local LaunchStats g_stats; void ffn_read_times(long long* out64); }
using namespace mesm;
#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 794624, iters = argc > 2 ? atoi(argv[2]) : 5;
    float *X, *R, *out, *W1, *W2, *v; void *w1f, *w2f;
    CKE(cudaMalloc(&X, (size_t)M * 256 * 4)); CKE(cudaMalloc(&R, (size_t)M * 256 * 4)); CKE(cudaMalloc(&out, (size_t)M * 256 * 4));
    CKE(cudaMalloc(&W1, 1024 * 256 * 4)); CKE(cudaMalloc(&W2, 1024 * 256 * 4)); CKE(cudaMalloc(&v, 4096 * 4));
    CKE(cudaMalloc(&w1f, ffn_packed_bytes())); CKE(cudaMalloc(&w2f, ffn_packed_bytes()));
    std::vector<float> h((size_t)1 << 20);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
    for (size_t o = 0; o < (size_t)M * 256; o += h.size()) { size_t n = std::min(h.size(), (size_t)M * 256 - o); CKE(cudaMemcpy(X + o, h.data(), n * 4, cudaMemcpyHostToDevice)); CKE(cudaMemcpy(R + o, h.data(), n * 4, cudaMemcpyHostToDevice)); }
    CKE(cudaMemcpy(W1, h.data(), 1024 * 256 * 4, cudaMemcpyHostToDevice)); CKE(cudaMemcpy(W2, h.data(), 1024 * 256 * 4, cudaMemcpyHostToDevice));
    CKE(cudaMemcpy(v, h.data(), 4096 * 4, cudaMemcpyHostToDevice));
    CKE(launch_pack_ffn(W1, W2, w1f, w2f, 0));
    FfnArgs a; a.X = X; a.ldx = 256; a.R = R; a.ldr = 256; a.out = out; a.ldo = 256; a.omap = identity_map(); a.M = M;
    a.W1f = w1f; a.W2f = w2f; a.maps = ffn_make_maps(w1f, w2f); a.b1 = v; a.b2 = v + 1024; a.ln_g = v + 1280; a.ln_b = v + 1536; a.prelu = v + 1792;
    if (!ffn_fused_eligible(a)) { printf("not eligible\n"); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CKE(launch_ffn_fused(a, 0)); CKE(cudaDeviceSynchronize());
    float tot = 0;
    for (int i = 0; i < iters; ++i) { cudaEventRecord(e0, 0); CKE(launch_ffn_fused(a, 0)); cudaEventRecord(e1, 0); CKE(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms; }
    const double fl = 4.0 * M * 256.0 * 1024.0;
    printf("M=%d: %.1f us -> %.1f TFLOP/s algorithmic (x3 issued %.1f)\n", M, tot / iters * 1e3, fl / (tot / iters * 1e-3) / 1e12, 3 * fl / (tot / iters * 1e-3) / 1e12);
#ifdef MESM_TC_TIMING
    long long t[160]; ffn_read_times(t);
    printf("setup %lld | first X block at MMA %lld | X converted %lld | yfull committed %lld | yfull seen %lld | done %lld\n", t[1] - t[0], t[2] - t[0], t[20] - t[0], t[19] - t[0], t[37] - t[0], t[38] - t[0]);
    printf("epilogue: yfull seen %lld | pass 1 (+b2 +R, sums) done %lld | pass 2 (variance) done %lld | pass 3 (normalise, store) done %lld\n", t[37] - t[0], t[150] - t[0], t[151] - t[0], t[38] - t[0]);
    printf("MMA thread waits (cycles): local slot full %lld | peer slot %lld | hacc_free %lld | hbf_full %lld\n", t[60], t[61], t[62], t[63]);
    for (int g = 0; g < 16; ++g)
        printf("  slot g=%d: producer issued %7lld | leader local full %7lld | peer relayed %7lld | commit issued %7lld\n", g + 20, t[64 + g] - t[0], t[80 + g] - t[0], t[96 + g] - t[0], t[112 + g] - t[0]);
    for (int j = 0; j < 8; ++j)
        printf("  chunk %d: G1 issued %7lld  G2 issued %7lld | E1 got Hacc %7lld math done %7lld Hbf free %7lld | producer seg G1 %7lld\n", j, t[3 + j] - t[0], t[11 + j] - t[0],
               t[21 + j] - t[0], t[56 + j] - t[0], t[29 + j] - t[0], t[40 + (j == 0 ? 0 : 2 * j - 1)] - t[0]);
#endif
    return 0;
}
