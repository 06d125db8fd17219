// Host replacements for the inline-PTX helpers of log_b200/csrc/lgr_blend.cu (tests/emu/build_emu.py removes the
// originals by name).  "Shared-window addresses" are 32-bit values: offsets from an anchor inside this library's data
// segment for static __shared__ arrays, 0x40000000 + offset into the launch's dynamic shared memory otherwise.
// ex2.approx / rcp.approx are replaced by exact exp2f / division: results agree with the GPU to the approximation
// error of those instructions (2 ulp), which the parity tolerance covers.  The tensor-core primitives (ldmatrix,
// mma.m16n8k8.tf32) are emulated lane by lane with the fragment layouts of the PTX ISA; the multiplier sees only the
// upper 19 bits of each fp32 operand, as the hardware does.
#pragma once
namespace lgr {
static char emu_shared_anchor;
inline uint32_t smem_u32(const void* p) {
  const char* dyn = reinterpret_cast<const char*>(emu::dyn_smem());
  const char* q = reinterpret_cast<const char*>(p);
  if (dyn && q >= dyn && q < dyn + (1 << 20)) return 0x40000000u + (uint32_t)(q - dyn);
  return (uint32_t)((uintptr_t)p - (uintptr_t)&emu_shared_anchor);
}
inline char* emu_shared_ptr(uint32_t a) {
  if (a >= 0x40000000u && a < 0x40000000u + (1u << 20)) return reinterpret_cast<char*>(emu::dyn_smem()) + (a - 0x40000000u);
  return &emu_shared_anchor + (int32_t)a;
}
inline float ex2_approx(float x) { return exp2f(x); }
inline float rcp_approx(float x) { return 1.0f / x; }
inline float4 lds_f4(uint32_t a) { return *reinterpret_cast<const float4*>(emu_shared_ptr(a)); }
inline float2 lds_f2(uint32_t a) { return *reinterpret_cast<const float2*>(emu_shared_ptr(a)); }
inline void sts_f32(uint32_t a, float v) { *reinterpret_cast<float*>(emu_shared_ptr(a)) = v; }
inline uint32_t pin_reg(uint32_t v) { return v; }
inline void red_shared_max_u32(uint32_t a, unsigned v) { atomicMax(reinterpret_cast<unsigned*>(emu_shared_ptr(a)), v); }
inline void red_shared_add_f32(uint32_t a, float v) { atomicAdd(reinterpret_cast<float*>(emu_shared_ptr(a)), v); }
inline void prefetch_l2(const float*) {}
// ldmatrix.sync.aligned.m8n8.x4.shared.b16: lane l supplies the address of row l%8 of block l/8; lane (g = l/4, t = l%4)
// receives the 32-bit word t of row g of each block
inline void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  const int lane = emu::lane_id(), g = lane >> 2, t = lane & 3;
  uint32_t* out[4] = {&r0, &r1, &r2, &r3};
  for (int b = 0; b < 4; b++) {
    const uint32_t row = __shfl_sync(0xffffffffu, addr, 8 * b + g);
    *out[b] = *reinterpret_cast<const uint32_t*>(emu_shared_ptr(row) + 4 * t);
  }
}
inline float emu_tf32(uint32_t bits) { return __uint_as_float(bits & 0xffffe000u); }
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (fragment layout: PTX ISA, "Matrix Fragments for mma.m16n8k8")
struct EmuFrag { uint32_t a[4], b[2]; };
inline void mma_tf32(float& d0, float& d1, float& d2, float& d3, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                     uint32_t b0, uint32_t b1) {
  const int lane = emu::lane_id(), g = lane >> 2, t = lane & 3;
  emu::publish<EmuFrag>(EmuFrag{{a0, a1, a2, a3}, {b0, b1}});      // one exchange round: every lane's six operand registers
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 8; k++) {
    const EmuFrag fa = emu::peek<EmuFrag>(4 * g + (k & 3));                    // A[g][k], A[g+8][k] live in lane (g, k%4)
    const float alo = emu_tf32(k < 4 ? fa.a[0] : fa.a[2]), ahi = emu_tf32(k < 4 ? fa.a[1] : fa.a[3]);
    const EmuFrag f0 = emu::peek<EmuFrag>(4 * (2 * t) + (k & 3)), f1 = emu::peek<EmuFrag>(4 * (2 * t + 1) + (k & 3));   // B[k][n]: lane (n, k%4)
    const float bn0 = emu_tf32(k < 4 ? f0.b[0] : f0.b[1]), bn1 = emu_tf32(k < 4 ? f1.b[0] : f1.b[1]);
    acc[0] += alo * bn0; acc[1] += alo * bn1; acc[2] += ahi * bn0; acc[3] += ahi * bn1;
  }
  emu::barrier_warp();
  d0 += acc[0]; d1 += acc[1]; d2 += acc[2]; d3 += acc[3];
}
}  // namespace lgr
