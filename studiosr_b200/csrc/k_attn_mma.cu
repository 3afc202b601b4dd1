// bf16 (shifted-)window attention for 8x8 windows on mma.sync tensor-core tiles with a
// register-resident softmax.  One CTA per window (64 tokens), 4 warps x 16 query rows, heads in a
// loop.  torch.roll / window_partition / window_reverse are pure addressing here: the CTA gathers
// its 64 source rows (cyclically shifted) with cp.async, and scatters the result back to the same
// rows.  Relative-position bias is looked up from a shared-memory copy of the table and the
// shift mask is derived from token coordinates, both applied to the score fragments in registers.
#include "ssr_device.cuh"

namespace ssr {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int AM_WS = 8;
constexpr int AM_N = 64;  // tokens per window

template <int DP>
__global__ void __launch_bounds__(128) attn_mma_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int row_bytes = a.ld_qkv * 2 + 16;  // +16 B: consecutive rows land in different banks for ldmatrix
  uint8_t* tile = smem;                     // [64][row_bytes]
  float* btab = reinterpret_cast<float*>(smem + AM_N * row_bytes);  // [heads][225]
  int* pix = reinterpret_cast<int*>(btab + a.heads * 225);          // [64]
  int* rid = pix + AM_N;                                            // [64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwx = a.W / AM_WS, nwy = a.H / AM_WS;
  int w = blockIdx.x;
  const int wx = w % nwx;
  w /= nwx;
  const int wy = w % nwy;
  const int b = w / nwy;

  if (tid < AM_N) {
    const int sy = wy * AM_WS + tid / AM_WS, sx = wx * AM_WS + tid % AM_WS;
    const int yy = (sy + a.shift) % a.H, xx = (sx + a.shift) % a.W;
    pix[tid] = (b * a.H + yy) * a.W + xx;
    rid[tid] = a.shift > 0 ? 3 * shift_region(sy, a.H, AM_WS, a.shift) + shift_region(sx, a.W, AM_WS, a.shift) : 0;
  }
  for (int e = tid; e < a.heads * 225; e += 128) btab[e] = __ldg(a.bias + e);
  __syncthreads();

  // gather the window: 64 rows x (3*QP) bf16
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv);
  const int chunks = a.ld_qkv / 8;  // 16-byte chunks per row
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  for (int e = tid; e < AM_N * chunks; e += 128) {
    const int r = e / chunks, c = e - r * chunks;
    cp_async16(tile_s + r * row_bytes + c * 16, qkv + (size_t)pix[r] * a.ld_qkv + c * 8);
  }
  cp_async_wait_all();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int i0 = warp * 16 + g, i1 = i0 + 8;  // the two query tokens this thread holds
  const int yi0 = i0 >> 3, xi0 = i0 & 7, yi1 = i1 >> 3, xi1 = i1 & 7;
  const int rid0 = rid[i0], rid1 = rid[i1];
  const bool masked = a.shift > 0;
  constexpr float LOG2E = 1.4426950408889634f;

  for (int h = 0; h < a.heads; ++h) {
    const int qoff = (h * DP) * 2, koff = (a.QP + h * DP) * 2, voff = (2 * a.QP + h * DP) * 2;  // byte offsets in a row
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(tile_s + (warp * 16 + (lane & 15)) * row_bytes + qoff + ks * 32 + (lane >> 4) * 16, a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // two key n-tiles per ldmatrix.x4
        const int mat = lane >> 3;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(tile_s + (np * 16 + (mat >> 1) * 8 + (lane & 7)) * row_bytes + koff + ks * 32 + (mat & 1) * 16, b0, b1, b2,
                b3);
        mma_bf16(s[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // bias + mask, row max
    const float* bt = btab + h * 225;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        const int yj = j >> 3, xj = j & 7;
        float v0 = s[nt][e] + bt[(yi0 - yj + 7) * 15 + (xi0 - xj + 7)];
        float v1 = s[nt][2 + e] + bt[(yi1 - yj + 7) * 15 + (xi1 - xj + 7)];
        if (masked) {
          const int rj = rid[j];
          if (rj != rid0) v0 += -100.0f;
          if (rj != rid1) v1 += -100.0f;
        }
        s[nt][e] = v0;
        s[nt][2 + e] = v1;
        m0 = fmaxf(m0, v0);
        m1 = fmaxf(m1, v1);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.0f, l1 = 0.0f;
    const float mb0 = m0 * LOG2E, mb1 = m1 * LOG2E;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = exp2f(fmaf(s[nt][e], LOG2E, -mb0));
        const float p1 = exp2f(fmaf(s[nt][2 + e], LOG2E, -mb1));
        s[nt][e] = p0;
        s[nt][2 + e] = p1;
        l0 += p0;
        l1 += p1;
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // O = P V
    float o[DP / 8][4];
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < DP / 16; ++np) {
        const int mat = lane >> 3;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(tile_s + (kk * 16 + (mat & 1) * 8 + (lane & 7)) * row_bytes + voff + np * 32 + (mat >> 1) * 16, b0, b1,
                  b2, b3);
        mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // normalise and park the head's output in this warp's (now dead) Q slots
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
      *reinterpret_cast<uint32_t*>(tile + i0 * row_bytes + qoff + (nt * 8 + 2 * t) * 2) =
          pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
      *reinterpret_cast<uint32_t*>(tile + i1 * row_bytes + qoff + (nt * 8 + 2 * t) * 2) =
          pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
  __syncthreads();
  // scatter [64][QP] back to the source rows (window_reverse + inverse roll)
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.o);
  const int ochunks = a.QP / 8;
  for (int e = tid; e < AM_N * ochunks; e += 128) {
    const int r = e / ochunks, c = e - r * ochunks;
    const uint4 v = *reinterpret_cast<const uint4*>(tile + r * row_bytes + c * 16);
    *reinterpret_cast<uint4*>(out + (size_t)pix[r] * a.ld_o + c * 8) = v;
  }
}

int launch_attn_mma(const AttnArgs& a, cudaStream_t s) {
  SSR_CHECK(a.ws == 8, SSR_E_INVALID, "attn_mma: only 8x8 windows (ws=%d)", a.ws);
  SSR_CHECK(a.DP == 16 || a.DP == 32, SSR_E_INVALID, "attn_mma: padded head dim %d not in {16,32}", a.DP);
  SSR_CHECK(a.H % 8 == 0 && a.W % 8 == 0, SSR_E_INVALID, "attn: %dx%d not a multiple of 8", a.H, a.W);
  SSR_CHECK(a.ld_qkv % 8 == 0 && a.QP % 8 == 0 && a.ld_o % 8 == 0, SSR_E_INVALID, "attn_mma: unaligned leading dims");
  const size_t smem = (size_t)AM_N * (a.ld_qkv * 2 + 16) + (size_t)a.heads * 225 * 4 + 2 * AM_N * 4;
  SSR_CHECK(smem <= 200 * 1024, SSR_E_INVALID, "attn_mma: window tile needs %zu B of shared memory", smem);
  const int nwin = a.B * (a.H / 8) * (a.W / 8);
  // algorithmic work: QK^T and PV, 2*64*64*d MACs... = 4*64*64*d FLOP per (window, head); q,k,v read + o written
  ProfScope prof("attn_mma", 4.0 * nwin * 64 * 64 * a.d * a.heads, 4.0 * nwin * 64 * a.heads * a.d * 2, s);
  static size_t attr16 = 0, attr32 = 0;
  if (a.DP == 32) {
    if (smem > attr32) {
      SSR_CUDA(cudaFuncSetAttribute(attn_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr32 = smem;
    }
    attn_mma_kernel<32><<<nwin, 128, smem, s>>>(a);
  } else {
    if (smem > attr16) {
      SSR_CUDA(cudaFuncSetAttribute(attn_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr16 = smem;
    }
    attn_mma_kernel<16><<<nwin, 128, smem, s>>>(a);
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
