// Per-tile front-to-back alpha compositing (forward) and the gradient sweep (backward).
//
// One CTA of 256 threads per 16x16 tile; warp w owns the 8x4 pixel sub-tile ((w&1)*8, (w>>1)*4), one pixel per
// lane.  The tile's depth-sorted list is staged through shared memory 256 splats at a time (coalesced id read,
// 3 x 16-byte gather per splat, 48-byte staged record).  Each warp then tests 32 staged splats at once against its
// sub-tile (lane = splat, conservative alpha>=1/255 box), ballots, and only walks the hits (lane = pixel, broadcast
// LDS).  With small splats this skips ~90 % of the (pixel, splat) pairs the classic per-thread loop evaluates.
// Backward: the same front-to-back walk (closed form of the published recurrence, see below); per hit every lane
// publishes two scalars and 27 lanes reduce them against fixed weights held in registers (pixel-coordinate moments
// and cotangent-weighted sums); the moments become gradients once per staged splat, then 3 vector atomics per splat.
#include "lgr_common.cuh"
#include "lgr_prof.cuh"

namespace lgr {

constexpr int BLEND_THREADS = TILE_PIX;   // 256
// minimum resident CTAs per SM the compiler must make room for (register cap = 65536 / (256 * N)); measured, see DESIGN.md
#ifndef LGR_FWD_MIN_CTAS
#define LGR_FWD_MIN_CTAS 5
#endif
#ifndef LGR_BWD_MIN_CTAS
#define LGR_BWD_MIN_CTAS 4
#endif
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
// kernel (the backward re-walks the list front to back and must stop where the forward stopped).  Explicitly rounded
// intrinsics are never contracted or re-associated, so both kernels execute the same arithmetic.
// The conic in the splat record is pre-multiplied by log2(e): alpha = o * 2^(power2).
__device__ __forceinline__ float eval_power2(const float4 r0, const float con_z, float dx, float dy) {
  const float q = __fmaf_rn(r0.z, __fmul_rn(dx, dx), __fmul_rn(con_z, __fmul_rn(dy, dy)));   // cx dx^2 + cz dy^2
  return __fmaf_rn(-0.5f, q, -__fmul_rn(r0.w, __fmul_rn(dx, dy)));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float eval_alpha(float opacity, float G) { return fminf(ALPHA_MAX, __fmul_rn(opacity, G)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
// keep a shared-window address in a register (the compiler otherwise rebuilds it from SR_CgaCtaId inside hot loops)
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ void red_shared_max_u32(uint32_t addr, unsigned v) {
  asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_add_f32(uint32_t addr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Staged splat: three consecutive float4 per list entry (48-byte stride: conflict-free for 128-bit accesses),
//   [0] = (px, py, conic_x', conic_y')   [1] = (conic_z', opacity, hx, hy)   [2] = (r, g, b, id as int bits)
__device__ __forceinline__ void stage_splat(float4* s_rec, int slot, const float* __restrict__ splat, int id) {
  const float* rec = splat + (int64_t)id * LGR_SPLAT_FLOATS;
  float4 r2 = ldg4(rec + 8);
  r2.w = __int_as_float(id);
  s_rec[3 * slot] = ldg4(rec); s_rec[3 * slot + 1] = ldg4(rec + 4); s_rec[3 * slot + 2] = r2;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <bool AUX>
__global__ void __launch_bounds__(BLEND_THREADS, LGR_FWD_MIN_CTAS)
blend_fwd_kernel(View v, const int32_t* __restrict__ tile_start, const int32_t* __restrict__ sorted_ids,
                 const float* __restrict__ splat, float* __restrict__ image, float* __restrict__ final_T,
                 int32_t* __restrict__ n_contrib, int32_t* __restrict__ pid_pixel, float* __restrict__ pw_pixel,
                 unsigned* __restrict__ point_weight_bits, int32_t* __restrict__ point_count) {
  __shared__ float4 s_rec[BATCH * 3];
  __shared__ unsigned s_w[AUX ? BATCH : 1];
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SubTile st = make_subtile(v, tile, lane, warp);
  const float pxf = (float)st.x, pyf = (float)st.y;
  const int beg = tile_start[tile], len = tile_start[tile + 1] - beg;
  const uint32_t s_w_addr = smem_u32(s_w);

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, wmax = 0.f;
  int last = 0, wid = -1;
  int done = st.inside ? 0 : 1;

  for (int base = 0; base < len; base += BATCH) {
    if (__syncthreads_and(done)) break;      // also orders smem reuse between batches
    const int cnt = min(BATCH, len - base);
    if (tid < cnt) {
      stage_splat(s_rec, tid, splat, sorted_ids[beg + base + tid]);
      if (AUX) s_w[tid] = 0u;
    }
    __syncthreads();
    if (!__all_sync(FULL, done)) {
      for (int c0 = 0; c0 < cnt; c0 += 32) {
        const int e_l = c0 + lane;
        bool hit = false;
        if (e_l < cnt) hit = box_hits(s_rec[3 * e_l], s_rec[3 * e_l + 1], st);
        unsigned mask = __ballot_sync(FULL, hit);
        unsigned own_w = 0u;                 // max weight of the splat this lane tested, over this warp's pixels
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const float4* rec = s_rec + 3 * (c0 + j);
          const float4 r0 = rec[0];
          const float2 r1 = *reinterpret_cast<const float2*>(rec + 1);    // (conic_z, opacity)
          const float dx = __fsub_rn(r0.x, pxf), dy = __fsub_rn(r0.y, pyf);
          const float power = eval_power2(r0, r1.x, dx, dy);
          const float alpha = eval_alpha(r1.y, ex2_approx(power));
          float w = 0.f;
          if (!done && power <= 0.0f && alpha >= ALPHA_MIN) {
            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            if (test_T < T_STOP) done = 1;
            else {
              w = alpha * T;
              const float4 r2 = rec[2];
              C0 = fmaf(r2.x, w, C0); C1 = fmaf(r2.y, w, C1); C2 = fmaf(r2.z, w, C2);
              T = test_T;
              last = base + c0 + j + 1;
              if (AUX && w > wmax) { wmax = w; wid = __float_as_int(r2.w); }
            }
          }
          if (AUX) {
            const unsigned m = __reduce_max_sync(FULL, __float_as_uint(w));   // w >= 0: uint order == float order
            if (lane == j) own_w = m;
          }
        }
        if (AUX && own_w) red_shared_max_u32(s_w_addr + 4u * e_l, own_w);
        if (__all_sync(FULL, done)) break;
      }
    }
    if (AUX) {
      __syncthreads();
      if (tid < cnt && s_w[tid]) atomicMax(point_weight_bits + __float_as_int(s_rec[3 * tid + 2].w), s_w[tid]);
    }
  }
  if (st.inside) {
    const int64_t pix = (int64_t)st.y * v.W + st.x, HW = (int64_t)v.H * v.W;
    image[pix] = C0 + T * __ldg(v.bg);
    image[HW + pix] = C1 + T * __ldg(v.bg + 1);
    image[2 * HW + pix] = C2 + T * __ldg(v.bg + 2);
    final_T[pix] = T;
    n_contrib[pix] = last;
    if (AUX) {
      pid_pixel[pix] = wid; pw_pixel[pix] = wmax;
      if (point_count && wid >= 0) atomicAdd(point_count + wid, 1);      // histogram of the per-pixel winners
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// The sweep runs FRONT TO BACK, exactly like the forward: T_j comes from the same multiplications (no division
// chain), the colour in front of j is the running prefix P_j, and the colour behind j follows from the rendered
// pixel:  S_j = image - P_j - c_j a_j T_j  (= sum_{k>j} c_k a_k T_k + bg T_final).  Then
//   dL/da_j = sum_c dL/dC_c * ( c_j T_j - S_j / (1 - a_j) )
// which is the published back-to-front recurrence written in closed form.
//
// Reduction over the warp's pixels.  Every one of the 9 per-splat outputs is a FIXED-weight linear functional of two
// per-lane scalars of the hit, wG = dL/dG * G and w = alpha * T:
//     M00, M10, M01, M20, M11, M02 = sum_l wG_l * {1, u, v, u^2, uv, v^2}_l      (u, v: tile-centred pixel coordinates)
//     C0, C1, C2                   = sum_l w_l * dL/dC_{0,1,2; l}
// so each lane publishes just (wG, w) to a per-warp shared scratch (two 4-byte stores, SoA), 27 lanes -- 9 outputs
// x 3 row groups -- each accumulate 12 rows (3 x LDS.128) against weights they keep in registers, the 3 partials are
// combined with two shuffles and 9 lanes add into the per-splat accumulators.  The moments are turned into
// d/dmean2D, d/dconic, d/dopacity once per staged splat when the batch is flushed (X = splat centre, same coordinates):
//     sum wG dx = X M00 - M10,   sum wG dx^2 = X^2 M00 - 2 X M10 + M20,   sum wG dx dy = XY M00 - X M01 - Y M10 + M11 ...
__global__ void __launch_bounds__(BLEND_THREADS, LGR_BWD_MIN_CTAS)
blend_bwd_kernel(View v, const int32_t* __restrict__ tile_start, const int32_t* __restrict__ sorted_ids,
                 const float* __restrict__ splat, const float* __restrict__ image,
                 const float* __restrict__ dL_dimage, float* __restrict__ dsplat) {
  __shared__ float4 s_rec[BATCH * 3];
  __shared__ float s_g[BATCH * 9];
  __shared__ __align__(16) float s_x[(BLEND_THREADS / 32) * 64];     // per warp: wG[32] | w[32]
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SubTile st = make_subtile(v, tile, lane, warp);
  const float pxf = (float)st.x, pyf = (float)st.y;
  const int beg = tile_start[tile], len = tile_start[tile + 1] - beg;
  float* xg = s_x + warp * 64;
  const uint32_t s_g_lane = pin_reg(smem_u32(s_g) + 4u * (uint32_t)lane);   // this lane's column of the accumulators
  const uint32_t s_rec_addr = pin_reg(smem_u32(s_rec));
  const float tcx = (float)((tile % v.gx) * TILE) + 7.5f, tcy = (float)((v.row0 + tile / v.gx) * TILE) + 7.5f;

  float I0 = 0.f, I1 = 0.f, I2 = 0.f, dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
  if (st.inside) {
    const int64_t pix = (int64_t)st.y * v.W + st.x, HW = (int64_t)v.H * v.W;
    I0 = image[pix]; I1 = image[HW + pix]; I2 = image[2 * HW + pix];
    dp0 = dL_dimage[pix]; dp1 = dL_dimage[HW + pix]; dp2 = dL_dimage[2 * HW + pix];
  }
  // fixed reduction weights of this lane: output red_k, rows 12 red_s .. 12 red_s + 11 (clipped to 32)
  const int red_k = lane % 9, red_s = min(lane / 9, 2);
  float wt[12];
  {
    const float u = pxf - tcx, w_ = pyf - tcy;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const int r = min(12 * red_s + i, 31);
      const float ur = __shfl_sync(FULL, u, r), vr = __shfl_sync(FULL, w_, r);
      const float d0 = __shfl_sync(FULL, dp0, r), d1 = __shfl_sync(FULL, dp1, r), d2 = __shfl_sync(FULL, dp2, r);
      float t;
      switch (red_k) {
        case 0: t = 1.f; break;
        case 1: t = ur; break;
        case 2: t = vr; break;
        case 3: t = ur * ur; break;
        case 4: t = ur * vr; break;
        case 5: t = vr * vr; break;
        case 6: t = d0; break;
        case 7: t = d1; break;
        default: t = d2; break;
      }
      wt[i] = (lane < 27 && 12 * red_s + i < 32) ? t : 0.f;
    }
  }
  // where this lane reads: the wG half for outputs 0..5, the w half for 6..8; group 2 re-reads quad 7 with weight 0
  const float4* red_src = reinterpret_cast<const float4*>(xg + (red_k >= 6 ? 32 : 0)) + 3 * red_s;
  const int q2 = red_s == 2 ? 1 : 2;

  float T = 1.0f, P0 = 0.f, P1 = 0.f, P2 = 0.f;
  int done = st.inside ? 0 : 1;

  for (int base = 0; base < len; base += BATCH) {
    if (__syncthreads_and(done)) break;
    const int cnt = min(BATCH, len - base);
    if (tid < cnt) stage_splat(s_rec, tid, splat, sorted_ids[beg + base + tid]);
    for (int k = tid; k < cnt * 9; k += BLEND_THREADS) s_g[k] = 0.f;
    __syncthreads();
    if (!__all_sync(FULL, done)) {
      for (int c0 = 0; c0 < cnt; c0 += 32) {
        const int e_l = c0 + lane;
        bool hit = false;
        if (e_l < cnt) hit = box_hits(s_rec[3 * e_l], s_rec[3 * e_l + 1], st);
        unsigned mask = __ballot_sync(FULL, hit);
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const int e = c0 + j;
          const uint32_t rec = s_rec_addr + 48u * (uint32_t)e;
          const float4 r0 = lds_f4(rec);
          const float2 r1 = lds_f2(rec + 16u);                            // (conic_z, opacity)
          const float dx = __fsub_rn(r0.x, pxf), dy = __fsub_rn(r0.y, pyf);
          const float power = eval_power2(r0, r1.x, dx, dy);
          const float G = ex2_approx(power);
          const float alpha = eval_alpha(r1.y, G);
          int contrib = 0;
          float test_T = 0.f;
          if (!done && power <= 0.0f && alpha >= ALPHA_MIN) {
            test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            if (test_T < T_STOP) done = 1; else contrib = 1;
          }
          if (!__any_sync(FULL, contrib)) continue;
          float wG = 0.f, w = 0.f;
          if (contrib) {
            const float4 r2 = lds_f4(rec + 32u);
            w = alpha * T;
            const float cdot = r2.x * dp0 + r2.y * dp1 + r2.z * dp2;
            // colour behind j (+ bg T_final):  S = I - P - c w
            const float Sdot = (I0 - P0) * dp0 + (I1 - P1) * dp1 + (I2 - P2) * dp2 - cdot * w;
            P0 = fmaf(r2.x, w, P0); P1 = fmaf(r2.y, w, P1); P2 = fmaf(r2.z, w, P2);
            const float dL_dalpha = cdot * T - Sdot * rcp_approx(1.0f - alpha);
            T = test_T;
            wG = r1.y * dL_dalpha * G;                      // dL/dG * G   (the 0.99 clamp is straight-through)
          }
          xg[lane] = wG; xg[32 + lane] = w;
          __syncwarp();
          const float4 a0 = red_src[0], a1 = red_src[1], a2 = red_src[q2];
          float sum = wt[0] * a0.x;
          sum = fmaf(wt[1], a0.y, sum); sum = fmaf(wt[2], a0.z, sum); sum = fmaf(wt[3], a0.w, sum);
          sum = fmaf(wt[4], a1.x, sum); sum = fmaf(wt[5], a1.y, sum); sum = fmaf(wt[6], a1.z, sum); sum = fmaf(wt[7], a1.w, sum);
          sum = fmaf(wt[8], a2.x, sum); sum = fmaf(wt[9], a2.y, sum); sum = fmaf(wt[10], a2.z, sum); sum = fmaf(wt[11], a2.w, sum);
          sum += __shfl_down_sync(FULL, sum, 9) + __shfl_down_sync(FULL, sum, 18);
          if (lane < 9) red_shared_add_f32(s_g_lane + 36u * (uint32_t)e, sum);
          __syncwarp();
        }
        if (__all_sync(FULL, done)) break;
      }
    }
    __syncthreads();
    if (tid < cnt) {
      const float* m = s_g + tid * 9;
      const float M00 = m[0], M10 = m[1], M01 = m[2], M20 = m[3], M11 = m[4], M02 = m[5];
      const bool nz = (M00 != 0.f) | (M10 != 0.f) | (M01 != 0.f) | (M20 != 0.f) | (M11 != 0.f) | (M02 != 0.f) |
                      (m[6] != 0.f) | (m[7] != 0.f) | (m[8] != 0.f);
      if (nz) {
        const float4 r0 = s_rec[3 * tid];
        const float2 r1 = *reinterpret_cast<const float2*>(&s_rec[3 * tid + 1]);
        const int id = __float_as_int(s_rec[3 * tid + 2].w);
        const float X = r0.x - tcx, Y = r0.y - tcy;
        const float Sx = fmaf(X, M00, -M10), Sy = fmaf(Y, M00, -M01);                       // sum wG dx, sum wG dy
        const float Sxx = fmaf(X, fmaf(X, M00, -2.f * M10), M20);                           // sum wG dx^2
        const float Syy = fmaf(Y, fmaf(Y, M00, -2.f * M01), M02);
        const float Sxy = fmaf(X, fmaf(Y, M00, -M01), fmaf(-Y, M10, M11));                  // sum wG dx dy
        float4 a, b;
        a.x = -(r0.z * Sx + r0.w * Sy);          // d/dpx  (x log2e: the conic in the record is pre-scaled)
        a.y = -(r1.x * Sy + r0.w * Sx);          // d/dpy  (x log2e)
        a.z = -0.5f * Sxx;                       // d/dconic_x
        a.w = -Sxy;                              // d/dconic_y
        b.x = -0.5f * Syy;                       // d/dconic_z
        b.y = M00 / r1.y;                        // d/dopacity = sum G dL/dalpha = sum wG / o
        b.z = m[6]; b.w = m[7];                  // d/drgb
        float4* dst = reinterpret_cast<float4*>(dsplat + (int64_t)id * LGR_GRAD_FLOATS);
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
        atomicAdd(reinterpret_cast<float*>(dst + 2), m[8]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
int launch_blend_fwd(const View& v, const int32_t* tile_start, const int32_t* sorted_ids, const float* splat,
                     float* image, float* final_T, int32_t* n_contrib, int32_t* pid_pixel, float* pw_pixel,
                     float* point_weight, int32_t* point_count, cudaStream_t st) {
  const int ntiles = v.gx * (v.row1 - v.row0);
  if (ntiles <= 0) return 0;
  ProfScope ps(K_BLEND_FWD, st);
  if (v.want_aux)
    blend_fwd_kernel<true><<<ntiles, BLEND_THREADS, 0, st>>>(v, tile_start, sorted_ids, splat, image, final_T, n_contrib, pid_pixel, pw_pixel, reinterpret_cast<unsigned*>(point_weight), point_count);
  else
    blend_fwd_kernel<false><<<ntiles, BLEND_THREADS, 0, st>>>(v, tile_start, sorted_ids, splat, image, final_T, n_contrib, pid_pixel, pw_pixel, reinterpret_cast<unsigned*>(point_weight), point_count);
  LGR_CHECK_LAUNCH();
  return 0;
}

int launch_blend_bwd(const View& v, const int32_t* tile_start, const int32_t* sorted_ids, const float* splat,
                     const float* image, const float* dL_dimage, float* dsplat, cudaStream_t st) {
  const int ntiles = v.gx * (v.row1 - v.row0);
  if (ntiles <= 0) return 0;
  ProfScope ps(K_BLEND_BWD, st);
  blend_bwd_kernel<<<ntiles, BLEND_THREADS, 0, st>>>(v, tile_start, sorted_ids, splat, image, dL_dimage, dsplat);
  LGR_CHECK_LAUNCH();
  return 0;
}

}  // namespace lgr
