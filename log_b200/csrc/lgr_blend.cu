// Per-tile front-to-back alpha compositing (forward) and the back-to-front gradient sweep (backward).
//
// One CTA of 256 threads per 16x16 tile; warp w owns the 8x4 pixel sub-tile ((w&1)*8, (w>>1)*4), one pixel per
// lane.  The tile's depth-sorted list is staged through shared memory 256 splats at a time (coalesced id read,
// 3 x 16-byte gather per splat).  Each warp then tests 32 staged splats at once against its sub-tile
// (lane = splat, conservative alpha>=1/255 box), ballots, and only walks the hits (lane = pixel, broadcast LDS).
// With small splats this skips ~3/4 of the (pixel, splat) pairs the classic per-thread loop evaluates.
// Backward: per hit the 9 partial gradients are reduced across the warp with a transposed butterfly (16 SHFL
// instead of 45), accumulated per staged splat in shared memory, and flushed with 3 vector atomics per splat.
#include "lgr_common.cuh"
#include "lgr_prof.cuh"

namespace lgr {

constexpr int BLEND_THREADS = TILE_PIX;   // 256
constexpr int BATCH = 256;
constexpr unsigned FULL = 0xffffffffu;

struct SubTile {
  int x, y;          // this lane's pixel
  bool inside;
  float x0, x1, y0, y1;   // pixel-centre bounds of the warp's 8x4 sub-tile (clipped to the image)
};

__device__ __forceinline__ SubTile make_subtile(const View& v, int tile, int lane, int warp) {
  SubTile s;
  const int tx = tile % v.gx, ty = v.row0 + tile / v.gx;
  const int sx = tx * TILE + (warp & 1) * 8, sy = ty * TILE + (warp >> 1) * 4;
  s.x = sx + (lane & 7); s.y = sy + (lane >> 3);
  s.inside = s.x < v.W && s.y < v.H;
  s.x0 = (float)sx; s.x1 = (float)min(sx + 7, v.W - 1);
  s.y0 = (float)sy; s.y1 = (float)min(sy + 3, v.H - 1);
  return s;
}

__device__ __forceinline__ bool box_hits(const float4 r0, const float4 r1, const SubTile& s) {
  return (r0.x + r1.z >= s.x0) && (r0.x - r1.z <= s.x1) && (r0.y + r1.w >= s.y0) && (r0.y - r1.w <= s.y1);
}

// The skip decisions (power > 0, alpha < 1/255, T < 1e-4) must come out IDENTICAL in the forward and the backward
// kernel, otherwise the transmittance the backward recovers by division drifts from the forward's.  Explicitly
// rounded intrinsics are never contracted or re-associated, so both kernels execute the same arithmetic.
__device__ __forceinline__ float eval_power(const float4 r0, const float4 r1, float dx, float dy) {
  const float q = __fmaf_rn(r0.z, __fmul_rn(dx, dx), __fmul_rn(r1.x, __fmul_rn(dy, dy)));   // cx dx^2 + cz dy^2
  return __fmaf_rn(-0.5f, q, -__fmul_rn(r0.w, __fmul_rn(dx, dy)));
}
__device__ __forceinline__ float eval_alpha(float opacity, float G) { return fminf(ALPHA_MAX, __fmul_rn(opacity, G)); }

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <bool AUX>
__global__ void __launch_bounds__(BLEND_THREADS)
blend_fwd_kernel(View v, const int32_t* __restrict__ tile_start, const int32_t* __restrict__ sorted_ids,
                 const float* __restrict__ splat, float* __restrict__ image, float* __restrict__ final_T,
                 int32_t* __restrict__ n_contrib, int32_t* __restrict__ pid_pixel, float* __restrict__ pw_pixel,
                 unsigned* __restrict__ point_weight_bits) {
  __shared__ float4 s_r0[BATCH], s_r1[BATCH], s_r2[BATCH];
  __shared__ int s_id[BATCH];
  __shared__ unsigned s_w[AUX ? BATCH : 1];
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SubTile st = make_subtile(v, tile, lane, warp);
  const float pxf = (float)st.x, pyf = (float)st.y;
  const int beg = tile_start[tile], len = tile_start[tile + 1] - beg;

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, wmax = 0.f;
  int last = 0, wid = -1;
  bool done = !st.inside;

  for (int base = 0; base < len; base += BATCH) {
    if (__syncthreads_and(done)) break;      // also orders smem reuse between batches
    const int cnt = min(BATCH, len - base);
    if (tid < cnt) {
      const int id = sorted_ids[beg + base + tid];
      const float* rec = splat + (int64_t)id * LGR_SPLAT_FLOATS;
      s_r0[tid] = ldg4(rec); s_r1[tid] = ldg4(rec + 4); s_r2[tid] = ldg4(rec + 8);
      s_id[tid] = id;
      if (AUX) s_w[tid] = 0u;
    }
    __syncthreads();
    if (!__all_sync(FULL, done)) {
      for (int c0 = 0; c0 < cnt; c0 += 32) {
        const int e_l = c0 + lane;
        bool hit = false;
        if (e_l < cnt) hit = box_hits(s_r0[e_l], s_r1[e_l], st);
        unsigned mask = __ballot_sync(FULL, hit);
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const int e = c0 + j;
          const float4 r0 = s_r0[e], r1 = s_r1[e];
          const float dx = __fsub_rn(r0.x, pxf), dy = __fsub_rn(r0.y, pyf);
          const float power = eval_power(r0, r1, dx, dy);
          const float alpha = eval_alpha(r1.y, __expf(power));
          bool contrib = !done && power <= 0.0f && alpha >= ALPHA_MIN;
          float w = 0.f;
          if (contrib) {
            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            if (test_T < T_STOP) { done = true; contrib = false; }
            else {
              w = alpha * T;
              const float4 r2 = s_r2[e];
              C0 += r2.x * w; C1 += r2.y * w; C2 += r2.z * w;
              T = test_T;
              last = base + e + 1;
              if (AUX && w > wmax) { wmax = w; wid = s_id[e]; }
            }
          }
          if (AUX) {
            const unsigned m = __reduce_max_sync(FULL, __float_as_uint(w));   // w >= 0: uint order == float order
            if (lane == 0 && m > s_w[e]) atomicMax(&s_w[e], m);
          }
        }
        if (__all_sync(FULL, done)) break;
      }
    }
    if (AUX) {
      __syncthreads();
      if (tid < cnt && s_w[tid]) atomicMax(point_weight_bits + s_id[tid], s_w[tid]);
    }
  }
  if (st.inside) {
    const int64_t pix = (int64_t)st.y * v.W + st.x, HW = (int64_t)v.H * v.W;
    image[pix] = C0 + T * __ldg(v.bg);
    image[HW + pix] = C1 + T * __ldg(v.bg + 1);
    image[2 * HW + pix] = C2 + T * __ldg(v.bg + 2);
    final_T[pix] = T;
    n_contrib[pix] = last;
    if (AUX) { pid_pixel[pix] = wid; pw_pixel[pix] = wmax; }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// Reduce 9 per-lane values over the warp.  On return lanes 2k and 2k+1 hold the warp total of slot k (k<16).
__device__ __forceinline__ float warp_reduce9_transposed(const float in[9], int lane) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 9; i++) v[i] = in[i];
#pragma unroll
  for (int i = 9; i < 16; i++) v[i] = 0.f;
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool hi = (lane & (half * 2)) != 0;
#pragma unroll
    for (int i = 0; i < half; i++) {
      const float a = v[i], b = v[i + half];
      const float send = hi ? a : b, keep = hi ? b : a;
      v[i] = keep + __shfl_xor_sync(FULL, send, half * 2);
    }
  }
  return v[0] + __shfl_xor_sync(FULL, v[0], 1);
}

__global__ void __launch_bounds__(BLEND_THREADS)
blend_bwd_kernel(View v, const int32_t* __restrict__ tile_start, const int32_t* __restrict__ sorted_ids,
                 const float* __restrict__ splat, const float* __restrict__ final_T,
                 const int32_t* __restrict__ n_contrib, const float* __restrict__ dL_dimage, float* __restrict__ dsplat) {
  __shared__ float4 s_r0[BATCH], s_r1[BATCH], s_r2[BATCH];
  __shared__ int s_id[BATCH];
  __shared__ float s_g[BATCH * 9];
  __shared__ int s_max;
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SubTile st = make_subtile(v, tile, lane, warp);
  const float pxf = (float)st.x, pyf = (float)st.y;
  const int beg = tile_start[tile];

  float T_final = 1.f, dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
  int ncon = 0;
  if (st.inside) {
    const int64_t pix = (int64_t)st.y * v.W + st.x, HW = (int64_t)v.H * v.W;
    T_final = final_T[pix]; ncon = n_contrib[pix];
    dp0 = dL_dimage[pix]; dp1 = dL_dimage[HW + pix]; dp2 = dL_dimage[2 * HW + pix];
  }
  const float bgdot = __ldg(v.bg) * dp0 + __ldg(v.bg + 1) * dp1 + __ldg(v.bg + 2) * dp2;
  if (tid == 0) s_max = 0;
  __syncthreads();
  {
    const int m = __reduce_max_sync(FULL, ncon);
    if (lane == 0 && m > 0) atomicMax(&s_max, m);
  }
  __syncthreads();
  const int maxc = s_max;
  const int wmaxc = __reduce_max_sync(FULL, ncon);   // this warp's deepest contributor

  float T = T_final, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;

  for (int base = ((maxc - 1) / BATCH) * BATCH; base >= 0 && maxc > 0; base -= BATCH) {
    const int cnt = min(BATCH, maxc - base);
    __syncthreads();
    if (tid < cnt) {
      const int id = sorted_ids[beg + base + tid];
      const float* rec = splat + (int64_t)id * LGR_SPLAT_FLOATS;
      s_r0[tid] = ldg4(rec); s_r1[tid] = ldg4(rec + 4); s_r2[tid] = ldg4(rec + 8);
      s_id[tid] = id;
    }
    for (int k = tid; k < cnt * 9; k += BLEND_THREADS) s_g[k] = 0.f;
    __syncthreads();
    if (base < wmaxc) {
      const int wcnt = min(cnt, wmaxc - base);
      for (int c0 = ((wcnt - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
        const int e_l = c0 + lane;
        bool hit = false;
        if (e_l < wcnt) hit = box_hits(s_r0[e_l], s_r1[e_l], st);
        unsigned mask = __ballot_sync(FULL, hit);
        while (mask) {
          const int j = 31 - __clz(mask);
          mask &= ~(1u << j);
          const int e = c0 + j;
          const float4 r0 = s_r0[e], r1 = s_r1[e];
          const float dx = __fsub_rn(r0.x, pxf), dy = __fsub_rn(r0.y, pyf);
          const float power = eval_power(r0, r1, dx, dy);
          const float G = __expf(power);
          const float alpha = eval_alpha(r1.y, G);
          const bool contrib = (base + e < ncon) && power <= 0.0f && alpha >= ALPHA_MIN;
          if (!__any_sync(FULL, contrib)) continue;
          float g[9];
#pragma unroll
          for (int k = 0; k < 9; k++) g[k] = 0.f;
          if (contrib) {
            const float4 r2 = s_r2[e];
            T = __fdiv_rn(T, __fsub_rn(1.0f, alpha));
            const float dch = alpha * T;
            acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
            acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
            acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
            lc0 = r2.x; lc1 = r2.y; lc2 = r2.z;
            float dL_dalpha = (r2.x - acc0) * dp0 + (r2.y - acc1) * dp1 + (r2.z - acc2) * dp2;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
            const float dL_dG = r1.y * dL_dalpha;      // 0.99 clamp is straight-through
            const float gdx = G * dx, gdy = G * dy;
            g[0] = dL_dG * (-gdx * r0.z - gdy * r0.w);
            g[1] = dL_dG * (-gdy * r1.x - gdx * r0.w);
            g[2] = -0.5f * gdx * dx * dL_dG;
            g[3] = -gdx * dy * dL_dG;
            g[4] = -0.5f * gdy * dy * dL_dG;
            g[5] = G * dL_dalpha;
            g[6] = dch * dp0; g[7] = dch * dp1; g[8] = dch * dp2;
          }
          const float tot = warp_reduce9_transposed(g, lane);
          if (!(lane & 1) && lane < 18) atomicAdd(&s_g[e * 9 + (lane >> 1)], tot);
        }
      }
    }
    __syncthreads();
    if (tid < cnt) {
      const float* gs = s_g + tid * 9;
      const float4 a = make_float4(gs[0], gs[1], gs[2], gs[3]);
      const float4 b = make_float4(gs[4], gs[5], gs[6], gs[7]);
      const float c = gs[8];
      const bool nz = (a.x != 0.f) | (a.y != 0.f) | (a.z != 0.f) | (a.w != 0.f) | (b.x != 0.f) | (b.y != 0.f) |
                      (b.z != 0.f) | (b.w != 0.f) | (c != 0.f);
      if (nz) {
        float4* dst = reinterpret_cast<float4*>(dsplat + (int64_t)s_id[tid] * LGR_GRAD_FLOATS);
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
        atomicAdd(reinterpret_cast<float*>(dst + 2), c);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
int launch_blend_fwd(const View& v, const int32_t* tile_start, const int32_t* sorted_ids, const float* splat,
                     float* image, float* final_T, int32_t* n_contrib, int32_t* pid_pixel, float* pw_pixel,
                     float* point_weight, cudaStream_t st) {
  const int ntiles = v.gx * (v.row1 - v.row0);
  if (ntiles <= 0) return 0;
  ProfScope ps(K_BLEND_FWD, st);
  if (v.want_aux)
    blend_fwd_kernel<true><<<ntiles, BLEND_THREADS, 0, st>>>(v, tile_start, sorted_ids, splat, image, final_T, n_contrib, pid_pixel, pw_pixel, reinterpret_cast<unsigned*>(point_weight));
  else
    blend_fwd_kernel<false><<<ntiles, BLEND_THREADS, 0, st>>>(v, tile_start, sorted_ids, splat, image, final_T, n_contrib, pid_pixel, pw_pixel, reinterpret_cast<unsigned*>(point_weight));
  LGR_CHECK_LAUNCH();
  return 0;
}

int launch_blend_bwd(const View& v, const int32_t* tile_start, const int32_t* sorted_ids, const float* splat,
                     const float* final_T, const int32_t* n_contrib, const float* dL_dimage, float* dsplat,
                     cudaStream_t st) {
  const int ntiles = v.gx * (v.row1 - v.row0);
  if (ntiles <= 0) return 0;
  ProfScope ps(K_BLEND_BWD, st);
  blend_bwd_kernel<<<ntiles, BLEND_THREADS, 0, st>>>(v, tile_start, sorted_ids, splat, final_T, n_contrib, dL_dimage, dsplat);
  LGR_CHECK_LAUNCH();
  return 0;
}

}  // namespace lgr
