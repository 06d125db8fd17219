// Host replacements for the inline-PTX helpers of log_b200/csrc/lgr_blend.cu (tests/emu/build_emu.py removes the
// originals by name).  "Shared-window addresses" are 32-bit offsets from an anchor inside this library's data segment.
// ex2.approx / rcp.approx are replaced by exact exp2f / division: results agree with the GPU to the approximation
// error of those instructions (2 ulp), which the parity tolerance covers.
#pragma once
namespace lgr {
static char emu_shared_anchor;
inline uint32_t smem_u32(const void* p) { return (uint32_t)((uintptr_t)p - (uintptr_t)&emu_shared_anchor); }
inline char* emu_shared_ptr(uint32_t a) { return &emu_shared_anchor + (int32_t)a; }
inline float ex2_approx(float x) { return exp2f(x); }
inline float rcp_approx(float x) { return 1.0f / x; }
inline float4 lds_f4(uint32_t a) { return *reinterpret_cast<const float4*>(emu_shared_ptr(a)); }
inline float2 lds_f2(uint32_t a) { return *reinterpret_cast<const float2*>(emu_shared_ptr(a)); }
inline uint32_t pin_reg(uint32_t v) { return v; }
inline void red_shared_max_u32(uint32_t a, unsigned v) { atomicMax(reinterpret_cast<unsigned*>(emu_shared_ptr(a)), v); }
inline void red_shared_add_f32(uint32_t a, float v) { atomicAdd(reinterpret_cast<float*>(emu_shared_ptr(a)), v); }
}  // namespace lgr
