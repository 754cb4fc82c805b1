// Timeline of fcomb_ts_kernel: clock64 stamps of CTA 0's issuer thread and of one epilogue warp per slot group
// (compiled with F2_TRACE).  Prints, per recorder, the event id and the cycles since the previous event.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I probabilistic-multiplanar-unet_b200/csrc \
//        -o /tmp/fcomb_trace scripts/fcomb_trace.cu && /tmp/fcomb_trace
// ids: epilogue 10+j wait acc / 20+j woke / 30+j ld done / 40+j st issued / 50+j st done / 60+j arrived (mid layers, slot j of
// the group); 70+j start of the head/L0 step, 80+j L0 arrived, 90+j softmax done; issuer 100+s wait ready / 110+s woke / 120+s issued.
#define F2_TRACE 1
#include "../probabilistic-multiplanar-unet_b200/csrc/api.cu"
#include "../probabilistic-multiplanar-unet_b200/csrc/fcomb_ts.cu"
#include <vector>
#include <cstdlib>

int main(int argc, char** argv) {
  const int B = (argc > 1) ? atoi(argv[1]) : 4, N = 16, L = 6, C = 3, nl = 4, H = 256, W = 256;
  const int64_t HW = (int64_t)H * W;
  void* feat; float *mu, *sigma, *eps, *w0, *b0, *wmid, *bmid, *wlast, *blast, *sums;
  cudaMalloc(&feat, B * HW * 64 * 2); cudaMemset(feat, 0x3c, B * HW * 64 * 2);
  auto fill = [](float** p, size_t n, float v) { cudaMalloc(p, n * 4); std::vector<float> h(n); for (size_t i = 0; i < n; ++i) h[i] = v * ((rand() % 2001) / 1000.f - 1.f); cudaMemcpy(*p, h.data(), n * 4, cudaMemcpyHostToDevice); };
  fill(&mu, B * L, 1.f); fill(&sigma, B * L, 1.f); fill(&eps, (size_t)B * N * L, 1.f); fill(&w0, 64 * (64 + L), 0.1f); fill(&b0, 64, 0.1f);
  fill(&wmid, 2 * 64 * 64, 0.1f); fill(&bmid, 2 * 64, 0.1f); fill(&wlast, C * 64, 0.1f); fill(&blast, C, 0.1f);
  cudaMalloc(&sums, (size_t)B * 2 * C * HW * 4);
  uint32_t* trace;
  cudaMalloc(&trace, 3 * 1024 * 4); cudaMemset(trace, 0, 3 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    uint32_t* t = (it == 2) ? trace : nullptr;
    cudaMemcpyToSymbol(pmu::f2_trace_buf, &t, sizeof(t));
    cudaEventRecord(e0);
    int rc = pmu_fcomb_softmax_accum_bf16(feat, mu, sigma, eps, w0, b0, wmid, bmid, wlast, blast, sums, B, N, L, C, nl, HW, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc %d %s %s\n", rc, pmu_last_error(), cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %.3f ms (B = %d slices of 256^2, N = %d) -> %.0f clk per tile-sample at 1.9 GHz\n", it, ms, B, N,
           ms * 1e-3 * 1.9e9 / ((double)B * 512 * N / 148));
  }
  std::vector<uint32_t> h(3 * 1024);
  cudaMemcpy(h.data(), trace, h.size() * 4, cudaMemcpyDeviceToHost);
  const char* names[3] = {"epilogue warp 0 (group 0)", "epilogue warp 8 (group 1)", "issuer"};
  for (int r = 0; r < 3; ++r) {
    printf("== %s ==\n", names[r]);
    uint32_t prev = 0; long long tabs = 0;
    int shown = 0;
    for (int i = 0; i < 1024 && shown < 320; ++i) {
      uint32_t w = h[r * 1024 + i];
      if (w == 0) break;
      uint32_t id = w >> 24, t = w & 0xFFFFFFu;
      uint32_t d = i ? ((t - prev) & 0xFFFFFFu) : 0;
      tabs += d;
      if (i >= 300) { printf("%3u@%lld(+%u) ", id, tabs, d); if (++shown % 8 == 0) printf("\n"); }
      prev = t;
    }
    printf("\n");
  }
  return 0;
}
