// dcn_box.cuh — the TMA-staged input box of the DCNv2 tensor-core kernels and its conflict-free gather order.
//
// A box is BH x BW pixels of one 8-channel chunk of the group-blocked input, [y][x][8 ch] fp32, loaded by one 3-D
// TMA copy (out-of-image parts arrive as zeros = the reference's per-corner bounds tests, im2col_cuda.cu:38-48). The
// pitch BW * 32 B = 960 B is 64 (mod 128): the four corners of a bilinear sample — 2 x 2 pixels x 32 B — are eight
// 16-byte chunks that tile all 32 shared-memory banks exactly once. Lane L visits its sample's chunks in the rotated
// order c = (c0 + i) mod 8 with c0 = (L - bank group of the first corner) mod 8, so the eight lanes of every
// quarter-warp sit on eight different bank groups whatever the offsets are: LDS.128 (and, for the backward's
// accumulation box, the same geometry for 32-bit atomics) without bank conflicts. Measured on B200
// (tools/microbench/scatter_probe.cu): 1.18 clk per sample against 3.54 in natural order.
#pragma once
#include <cstdint>

namespace ebfi_dcn {
namespace box {

constexpr int BH = 24, BW = 30;                 // rows x pixels (8 channels each)
constexpr int PITCH = BW * 32;                  // bytes
constexpr int BYTES = BH * PITCH;               // 23,040 (a multiple of 128)
static_assert(PITCH % 128 == 64, "the 2x2 corner neighbourhood must cover all 32 banks");

__device__ __forceinline__ float4 lds128(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// Packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2): one issue slot for two lanes of math. The kernels that use the box
// are bound by instruction issue, not by the FMA pipe, so halving the FMA instruction count is a direct gain. A scalar
// operand is written as pack2(s, s): ptxas folds it into the instruction's broadcast form (R.F32), no extra move.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Visiting order of one in-box sample. Chunk c (0..7): corner c >> 1 (0 (y0,x0), 1 (y0,x0+1), 2 (y0+1,x0), 3 (y0+1,x0+1)),
// 16-byte half c & 1 (channels 0-3 / 4-7). Step i reads chunk (c0 + i) & 7: even steps are one half, odd steps the
// other — half 0 first iff !odd.
struct Rot {
    uint32_t base;       // shared address of corner (y0, x0)
    uint32_t r_lo, r_hi; // chunk offsets / 16 of steps 0..3 / 4..7, one byte each
    uint32_t c0;
    bool odd;
};

__device__ __forceinline__ Rot make_rot(uint32_t box_s, int yb, int xb, int lane)
{
    Rot r;
    r.base = box_s + (uint32_t)(yb * BW + xb) * 32u;
    r.c0 = ((uint32_t)lane - (r.base >> 4)) & 7u;
    r.odd = r.c0 & 1u;
    // byte table of chunk offsets / 16 = {0,1,2,3, P,P+1,P+2,P+3} with P = PITCH / 16, rotated by c0 bytes
    constexpr uint32_t T_LO = 0x03020100u, T_HI = 0x03020100u + 0x01010101u * (PITCH / 16);
    const uint32_t ta = (r.c0 & 4u) ? T_HI : T_LO, tb = (r.c0 & 4u) ? T_LO : T_HI, sh = (r.c0 & 3u) * 8u;
    r.r_lo = __funnelshift_r(ta, tb, sh);
    r.r_hi = __funnelshift_r(tb, ta, sh);
    return r;
}

// shared address of the chunk visited at (compile-time) step i
__device__ __forceinline__ uint32_t step_addr(const Rot &r, int i)
{
    return r.base + (__byte_perm(i < 4 ? r.r_lo : r.r_hi, 0u, 0x4440u | (uint32_t)(i & 3)) << 4);
}

// rotate per-corner coefficients so that w[j] belongs to corner (c0 / 2 + j) mod 4; w[4] = w[0]
__device__ __forceinline__ void rot_corners(uint32_t c0, float w0, float w1, float w2, float w3, float (&w)[5])
{
    if (c0 & 2u) { const float t = w0; w0 = w1; w1 = w2; w2 = w3; w3 = t; }
    if (c0 & 4u) { float t = w0; w0 = w2; w2 = t; t = w1; w1 = w3; w3 = t; }
    w[0] = w0; w[1] = w1; w[2] = w2; w[3] = w3; w[4] = w0;
}

// coefficient of (compile-time) step i from a rotated array
__device__ __forceinline__ float step_coef(const float (&w)[5], bool odd, int i) { return odd ? w[(i + 1) >> 1] : w[i >> 1]; }

}  // namespace box
}  // namespace ebfi_dcn
