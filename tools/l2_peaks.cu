// L2 micro-benchmarks for the hash-grid kernels' real roofline: the gathers /
// reductions of the fine hash-grid levels are random 8-byte accesses into a 4 MB table that lives in L2, so
// HBM bandwidth is not what bounds them.  This measures, on the box it runs on,
//   (1) random 8-byte GATHER rate from an L2-resident table  (one row per lane, and lane PAIRS in one line)
//   (2) random 8-byte REDUCTION rate (RED.ADD.F32x2) into an L2-resident table (same two patterns), and
//       RED.ADD.F32x4 for pairs sharing a 16-byte slot
// and prints one JSON object.  Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_peaks
// tools/l2_peaks.cu && ./l2_peaks     (tools/l2_peaks.sh does both and stores profiles/round2_l2_peaks.json)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {      // cheap integer hash (avalanche)
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// MODE 0: every lane its own random row; MODE 1: lanes 2k / 2k+1 take rows r and r ^ 1 (one 16-byte slot);
// MODE 2: lanes 2k / 2k+1 take two rows of one 128-byte line (r and r ^ 8)
template <int MODE>
__device__ __forceinline__ uint32_t pick_row(uint32_t item, uint32_t mask) {
  if (MODE == 0) return mix(item) & mask;
  if (MODE == 3) {
    // a bundle of the grouped kernels at M = 4: 8 adjacent lanes = 4 groups x 2 dim-0 corners of one
    // pixel corner under the general linear layout -- two groups with both rows in one 16-byte slot, two whose second
    // row lies 16 rows further: 8 rows in 2 lines / 5 sectors
    const uint32_t ofs[8] = {0u, 1u, 2u, 3u, 4u, 20u, 8u, 24u};
    return (mix(item >> 3) & mask & ~31u) ^ ofs[item & 7u];
  }
  const uint32_t base = mix(item >> 1) & mask;
  return MODE == 1 ? (base ^ (item & 1u)) : (base ^ ((item & 1u) << 3));
}

template <int MODE, int U>
__global__ void __launch_bounds__(256) gather_kernel(const float2* __restrict__ table, uint32_t mask, uint32_t n_items,
                                                      float2* __restrict__ sink) {
  float2 acc = make_float2(0.f, 0.f);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride * U) {
    float2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(table + pick_row<MODE>(i + u * stride, mask));
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
  }
  if (acc.x == 12345.678f) sink[0] = acc;     // keeps the loads alive
}

template <int MODE>
__global__ void __launch_bounds__(256) red_kernel(float2* __restrict__ table, uint32_t mask, uint32_t n_items) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride)
    atomicAdd(table + pick_row<MODE>(i, mask), make_float2(1.0f, 0.5f));
}

// pairs merged into one 128-bit reduction issued by the even lane (what hashgrid_bwd_pair_kernel does when
// the two rows share a 16-byte slot)
__global__ void __launch_bounds__(256) red4_kernel(float2* __restrict__ table, uint32_t mask, uint32_t n_items) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
    if (i & 1u) continue;
    const uint32_t r = (mix(i >> 1) & mask) & ~1u;
    atomicAdd(reinterpret_cast<float4*>(table + r), make_float4(1.0f, 0.5f, 1.0f, 0.5f));
  }
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const uint32_t rows = 1u << 19;                 // one hash-grid level: 2^19 rows x 8 bytes = 4 MB
  const uint32_t n_items = 1u << 26;              // 67 M accesses per launch
  float2* table = nullptr;
  float2* sink = nullptr;
  cudaMalloc(&table, (size_t)rows * sizeof(float2));
  cudaMalloc(&sink, 64);
  cudaMemset(table, 0, (size_t)rows * sizeof(float2));
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8;
  const uint32_t mask = rows - 1;
  const float g0 = time_ms([&] { gather_kernel<0, 8><<<grid, 256>>>(table, mask, n_items, sink); }, 5);
  const float g1 = time_ms([&] { gather_kernel<1, 8><<<grid, 256>>>(table, mask, n_items, sink); }, 5);
  const float g2 = time_ms([&] { gather_kernel<2, 8><<<grid, 256>>>(table, mask, n_items, sink); }, 5);
  const float g3 = time_ms([&] { gather_kernel<3, 8><<<grid, 256>>>(table, mask, n_items, sink); }, 5);
  const float r0 = time_ms([&] { red_kernel<0><<<grid, 256>>>(table, mask, n_items); }, 5);
  const float r1 = time_ms([&] { red_kernel<1><<<grid, 256>>>(table, mask, n_items); }, 5);
  const float r2 = time_ms([&] { red_kernel<2><<<grid, 256>>>(table, mask, n_items); }, 5);
  const float r4 = time_ms([&] { red4_kernel<<<grid, 256>>>(table, mask, n_items); }, 5);
  const double n = (double)n_items;
  printf("{\"table_bytes\": %u, \"accesses_per_launch\": %u, \"sms\": %d,\n", rows * 8u, n_items, sms);
  printf(" \"gather_8B_random_Gps\": %.2f, \"gather_8B_pair_same_slot_Gps\": %.2f, \"gather_8B_pair_same_line_Gps\": %.2f,\n",
         n / g0 * 1e-6, n / g1 * 1e-6, n / g2 * 1e-6);
  printf(" \"gather_8B_bundle8_two_lines_Gps\": %.2f,\n", n / g3 * 1e-6);
  printf(" \"red_f32x2_random_Gps\": %.2f, \"red_f32x2_pair_same_slot_Gps\": %.2f, \"red_f32x2_pair_same_line_Gps\": %.2f,\n",
         n / r0 * 1e-6, n / r1 * 1e-6, n / r2 * 1e-6);
  printf(" \"red_f32x4_merged_pairs_rows_Gps\": %.2f,\n", n / r4 * 1e-6);
  printf(" \"note\": \"G accesses (table rows) per second; one access = 8 bytes = one 32-byte L2 sector touched\"}\n");
  cudaFree(table);
  cudaFree(sink);
  return 0;
}
