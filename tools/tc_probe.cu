// Developer probe (not part of the product library): times one fused-linear tcgen05 launch shape in isolation and dumps the
// in-kernel clock64 stamps of CTA (0,0) (build with -DMESM_TC_TIMING).  Usage: tc_probe M N K [pos] [res_ln] [iters] [pair]
#include "../mesm_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mesm {
thread_local LaunchStats g_stats;
void tc_read_times(long long* out64);
}
using namespace mesm;

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 74496, N = argc > 2 ? atoi(argv[2]) : 256, K = argc > 3 ? atoi(argv[3]) : 256;
    const int pos = argc > 4 ? atoi(argv[4]) : 0, resln = argc > 5 ? atoi(argv[5]) : 0, iters = argc > 6 ? atoi(argv[6]) : 20;
    const int pair = argc > 7 ? atoi(argv[7]) : 1;
    const int flush = argc > 8 ? atoi(argv[8]) : 1, flags = argc > 9 ? atoi(argv[9]) : 0;
    float *A, *P, *W, *out, *R, *g, *b, *bias; void* Wp;
    CKE(cudaMalloc(&A, (size_t)M * K * 4)); CKE(cudaMalloc(&P, (size_t)M * K * 4)); CKE(cudaMalloc(&W, (size_t)N * K * 4));
    CKE(cudaMalloc(&out, (size_t)M * N * 4)); CKE(cudaMalloc(&R, (size_t)M * N * 4)); CKE(cudaMalloc(&g, N * 4)); CKE(cudaMalloc(&b, N * 4));
    CKE(cudaMalloc(&bias, N * 4)); CKE(cudaMalloc(&Wp, tc_packed_bytes(N, K)));
    std::vector<float> h((size_t)M * K);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
    CKE(cudaMemcpy(A, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); CKE(cudaMemcpy(P, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CKE(cudaMemcpy(W, h.data(), (size_t)N * K * 4, cudaMemcpyHostToDevice));
    CKE(cudaMemset(R, 0, (size_t)M * N * 4)); CKE(cudaMemcpy(g, h.data(), N * 4, cudaMemcpyHostToDevice));
    CKE(cudaMemcpy(b, h.data(), N * 4, cudaMemcpyHostToDevice)); CKE(cudaMemcpy(bias, h.data(), N * 4, cudaMemcpyHostToDevice));
    CKE(launch_pack_tc(W, 0, N, K, nullptr, Wp, 0));
    LinearOp op = make_linear(M, N, K, A, K, nullptr, (N + 3) / 4 * 4, bias, out, N);
    op.Wp = Wp; op.Apos = pos ? P : nullptr;
    if (resln) { op.residual = R; op.ldr = N; op.ln_g = g; op.ln_b = b; }
    if (!linear_tc_eligible(op)) { printf("not eligible\n"); return 1; }
    // a second, cache-thrashing buffer set so that consecutive launches do not find their inputs in L2
    float* big; const size_t bigN = (size_t)96 << 20; CKE(cudaMalloc(&big, bigN * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) CKE(launch_linear_tc(op, 0));
    CKE(cudaDeviceSynchronize());
    float tot = 0, best = 1e9;
    for (int i = 0; i < iters; ++i) {
        if (flush) CKE(cudaMemsetAsync(big, i, bigN * 4, 0));            // flush L2 (384 MB written)
        cudaEventRecord(e0, 0);
        CKE(launch_linear_tc(op, 0));
        cudaEventRecord(e1, 0);
        CKE(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms; best = ms < best ? ms : best;
    }
    const double fl = 2.0 * M * N * (double)K * (pos ? 2 : 1);
    printf("flush=%d flags=%d ", flush, flags);
    printf("M=%d N=%d K=%d pos=%d resln=%d pair=%d : avg %.1f us best %.1f us  -> %.1f TFLOP/s algorithmic (x3 issued: %.1f)\n", M, N, K, pos,
           resln, pair, tot / iters * 1e3, best * 1e3, fl / (tot / iters * 1e-3) / 1e12, 3 * fl / (tot / iters * 1e-3) / 1e12);
#ifdef MESM_TC_TIMING
    long long t[64]; tc_read_times(t);
    printf("stamps rel. to start (cycles): setup %lld | mma-issue-done %lld | epi-start %lld | epi-end %lld | exit %lld\n", t[1] - t[0],
           t[2] - t[0], t[3] - t[0], t[4] - t[0], t[6] - t[0]);
    for (int kb = 0; kb < 8; ++kb)
        printf("  kb%d: mma wfull %lld afull %lld | conv before-empty %lld after-empty %lld arrived %lld\n", kb, t[8 + kb] - t[0],
               t[16 + kb] - t[0], t[24 + kb] - t[0], t[32 + kb] - t[0], t[40 + kb] - t[0]);
    printf("  epi: c0 %lld %lld %lld %lld | c1 %lld %lld %lld %lld\n", t[48] - t[0], t[49] - t[0], t[50] - t[0], t[51] - t[0], t[52] - t[0],
           t[53] - t[0], t[54] - t[0], t[55] - t[0]);
#endif
    return 0;
}
