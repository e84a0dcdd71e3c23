// lgk_heights.cu -- terrain height field helpers (reference LR:831-869, MATH:38-42).
#include "lgk_math.cuh"

namespace lgk {

// min3[r,c] = min(hs[r,c], hs[r+1,c], hs[r,c+1]) -- the three samples LR:863-867 takes.  Because the
// reference clips the indices to r<=rows-2, c<=cols-2 BEFORE sampling (LR:860-861), one gather from this
// field returns exactly what three gathers + two mins return there.  Integer work: bit-exact.
__global__ void height_min3_kernel(const int16_t* __restrict__ hs, int16_t* __restrict__ out, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    int16_t v = hs[i];
    if (r + 1 < rows && c + 1 < cols) {
      const int16_t a = hs[i + cols], b = hs[i + 1];
      v = v < a ? v : a;
      v = v < b ? v : b;
    }
    out[i] = v;
  }
}

// stand-alone _get_heights: one warp per env, lanes over points
__global__ void __launch_bounds__(128) height_scan_kernel(const float* __restrict__ root_states, int actors_per_env,
                                                         int root_actor_offset, int num_envs,
                                                         const float* __restrict__ pts, int P,
                                                         const int16_t* __restrict__ min3, int rows, int cols,
                                                         float border, float hscale, float vscale,
                                                         float* __restrict__ out, int* __restrict__ px_out,
                                                         int* __restrict__ py_out) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  const float* r = root_states + ((size_t)env * actors_per_env + root_actor_offset) * 13;
  const YawFrame f = yaw_frame(r[5], r[6], r[0], r[1]);
  for (int j = lane; j < P; j += 32) {
    int ix, iy;
    height_index(f, pts[2 * j], pts[2 * j + 1], border, hscale, rows, cols, ix, iy);
    const int16_t h = min3[(size_t)ix * cols + iy];
    out[(size_t)env * P + j] = f_mul((float)h, vscale);
    if (px_out) px_out[(size_t)env * P + j] = ix;
    if (py_out) py_out[(size_t)env * P + j] = iy;
  }
}

}  // namespace lgk

using namespace lgk;

extern "C" int lgk_height_min3(const int16_t* height_samples, int16_t* min3, int32_t rows, int32_t cols, void* stream) {
  LGK_REQUIRE(height_samples && min3 && rows >= 2 && cols >= 2, "height_min3: bad arguments");
  height_min3_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(height_samples, min3, rows, cols);
  count_launch();
  return check_cuda(cudaGetLastError(), "height_min3_kernel launch");
}

extern "C" int lgk_height_scan(const float* root_states, int32_t actors_per_env, int32_t root_actor_offset,
                               int32_t num_envs, const float* height_points_xy, int32_t num_points,
                               const int16_t* height_min3, int32_t rows, int32_t cols, float border_size,
                               float horizontal_scale, float vertical_scale, float* measured_heights,
                               int32_t* px_out, int32_t* py_out, void* stream) {
  LGK_REQUIRE(root_states && height_points_xy && height_min3 && measured_heights, "height_scan: null pointer");
  LGK_REQUIRE(num_envs > 0 && num_points > 0 && rows >= 2 && cols >= 2 && actors_per_env >= 1, "height_scan: bad sizes");
  height_scan_kernel<<<(num_envs + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      root_states, actors_per_env, root_actor_offset, num_envs, height_points_xy, num_points, height_min3, rows, cols,
      border_size, horizontal_scale, vertical_scale, measured_heights, px_out, py_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "height_scan_kernel launch");
}
