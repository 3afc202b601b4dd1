// bf16 window attention with an online softmax over key chunks (mma.sync tiles) for the large windows of HAT:
//   * (shifted-)window self-attention, 16x16 windows = 256 tokens (hat.py:84-111 inside HAB :154-195), and
//   * overlapping cross-attention: queries = a 16x16 window, keys / values = the 24x24 window around it cut out by
//     nn.Unfold(kernel 24, stride 16, padding 4) -- out-of-image keys are ZERO vectors that still take part in the
//     softmax with their bias -- with the reference's negative-index wrap of the bias lookup (hat.py:257-284, 490-513).
// One CTA = one (window, head): K and V of the window (Nk x 32 bf16 each) are gathered once into shared memory, then
// the 4 warps walk the queries in blocks of 64 (16 rows per warp), each block streaming the keys in chunks of 64:
// S = q k^T on mma.sync, bias / mask and the running max / sum in registers, P.V accumulated in registers.
// torch.roll / window_partition / unfold / window_reverse are addressing (hat.py:170-187, 251-263).
#include "ssr_device.cuh"

namespace ssr {

namespace {
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
__device__ __forceinline__ float ex2_fast(float x) {  // MUFU.EX2 only (exp2f adds range handling around it)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
}  // namespace

// oca = 0: keys = the query window itself (roll by `shift`, region mask); oca = 1: keys = kws x kws unfold window
template <int DP>
__global__ void __launch_bounds__(128) attn_flash_kernel(const AttnArgs a, const int oca) {
  constexpr int RB = DP * 2 + 16;
  constexpr int CH = DP / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int Nq = a.ws * a.ws, kws = oca ? a.kws : a.ws, Nk = kws * kws;
  const int nb = a.ws + kws - 1;
  uint8_t* tK = smem;
  uint8_t* tV = tK + (size_t)Nk * RB;
  uint8_t* tQ = tV + (size_t)Nk * RB;                       // [64][RB] queries, later the output staging
  float* btab = reinterpret_cast<float*>(tQ + 64 * RB);     // [nb*nb]
  int* kpix = reinterpret_cast<int*>(btab + nb * nb);       // [Nk] source row or -1
  int* kinfo = kpix + Nk;                                   // [Nk] (ky*nb + kx) | region << 16
  int* qpix = kinfo + Nk;                                   // [Nq]
  int* qinfo = qpix + Nq;                                   // [Nq]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y;
  const int nwx = a.W / a.ws, nwy = a.H / a.ws;
  int w = blockIdx.x;
  const int wx = w % nwx;
  w /= nwx;
  const int wy = w % nwy, b = w / nwy;
  const int shift = oca ? 0 : a.shift;
  const int rid_first = shift > 0 ? 3 * shift_region(wy * a.ws, a.H, a.ws, shift) + shift_region(wx * a.ws, a.W, a.ws, shift) : 0;
  int any_mixed = 0;  // does this window straddle shift-mask regions?  (only the last window row / column does)
  for (int t = tid; t < Nq; t += 128) {
    const int qy = t / a.ws, qx = t % a.ws;
    const int sy = wy * a.ws + qy, sx = wx * a.ws + qx;
    const int yy = (sy + shift) % a.H, xx = (sx + shift) % a.W;
    qpix[t] = (b * a.H + yy) * a.W + xx;
    const int rid = shift > 0 ? 3 * shift_region(sy, a.H, a.ws, shift) + shift_region(sx, a.W, a.ws, shift) : 0;
    any_mixed |= rid != rid_first;
    qinfo[t] = (qy * nb + qx) | (rid << 16);
    if (!oca) {
      kpix[t] = qpix[t];
      kinfo[t] = qinfo[t];
    }
  }
  if (oca) {
    const int pad = (kws - a.ws) / 2;
    for (int t = tid; t < Nk; t += 128) {
      const int ky = t / kws, kx = t % kws;
      const int y = wy * a.ws - pad + ky, x = wx * a.ws - pad + kx;
      kpix[t] = (y >= 0 && y < a.H && x >= 0 && x < a.W) ? (b * a.H + y) * a.W + x : -1;
      kinfo[t] = ky * nb + kx;
    }
  }
  for (int e = tid; e < nb * nb; e += 128) btab[e] = __ldg(a.bias + (size_t)h * nb * nb + e);
  const bool mixed = __syncthreads_or(any_mixed) != 0;

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv);
  const uint32_t sK = (uint32_t)__cvta_generic_to_shared(tK), sV = (uint32_t)__cvta_generic_to_shared(tV);
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(tQ);
  for (int e = tid; e < Nk * CH * 2; e += 128) {
    const int part = e / (Nk * CH), r = (e / CH) % Nk, c = e % CH;
    const uint32_t dst = (part ? sV : sK) + r * RB + c * 16;
    if (kpix[r] >= 0)
      cp_async16(dst, qkv + (size_t)kpix[r] * a.ld_qkv + (1 + part) * a.QP + h * DP + c * 8);
    else
      *reinterpret_cast<uint4*>((part ? tV : tK) + r * RB + c * 16) = make_uint4(0, 0, 0, 0);
  }

  const int g = lane >> 2, t4 = lane & 3;
  const bool masked = shift > 0 && mixed;  // a window inside one region needs no mask at all
  // 16x16 self-attention: key j of a chunk sits at (ky, kx) = (kc / 16 + nt / 2, (nt & 1) * 8 + 2 t + e), so the bias index
  // q - k is a per-chunk base minus a compile-time offset: one LDS with an immediate per score instead of two LDS + index math
  const bool fast16 = !oca && a.ws == 16;
  constexpr float LOG2E = 1.4426950408889634f;
  // bias index: self-attention (q - k + ws - 1) * nb + ..., OCA (k - q + ws - kws + 1) * nb + ... wrapped when negative
  const int qoff = oca ? -(a.ws - kws + 1) * (nb + 1) : (a.ws - 1) * (nb + 1);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.o);

  for (int qb = 0; qb < Nq; qb += 64) {
    __syncthreads();  // the previous block's output scatter has left tQ
    for (int e = tid; e < 64 * CH; e += 128) {
      const int r = e / CH, c = e % CH;
      cp_async16(sQ + r * RB + c * 16, qkv + (size_t)qpix[qb + r] * a.ld_qkv + h * DP + c * 8);
    }
    cp_async_wait_all();
    __syncthreads();
    uint32_t qf[DP / 16][4];
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks)
      ldsm_x4(sQ + (warp * 16 + (lane & 15)) * RB + ks * 32 + (lane >> 4) * 16, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    const int i0 = qb + warp * 16 + g, i1 = i0 + 8;
    const int qi0 = qinfo[i0], qi1 = qinfo[i1];
    const int qb0 = (qi0 & 0xffff) + qoff, qb1 = (qi1 & 0xffff) + qoff;
    const int rid0 = qi0 >> 16, rid1 = qi1 >> 16;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    float o[DP / 8][4];
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;

    for (int kc = 0; kc < Nk; kc += 64) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < DP / 16; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const int mat = lane >> 3;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(sK + (kc + np * 16 + (mat >> 1) * 8 + (lane & 7)) * RB + ks * 32 + (mat & 1) * 16, b0, b1, b2, b3);
          mma_bf16(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
          mma_bf16(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
        }
      }
      float c0 = -INFINITY, c1 = -INFINITY;
      if (fast16 && !masked) {
        const float* bt0 = btab + (qb0 - (kc >> 4) * 31 - 2 * t4);
        const float* bt1 = btab + (qb1 - (kc >> 4) * 31 - 2 * t4);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int off = (nt >> 1) * 31 + (nt & 1) * 8 + e;
            const float v0 = s[nt][e] + bt0[-off], v1 = s[nt][2 + e] + bt1[-off];
            s[nt][e] = v0;
            s[nt][2 + e] = v1;
            c0 = fmaxf(c0, v0);
            c1 = fmaxf(c1, v1);
          }
        }
      } else
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ki = kinfo[kc + nt * 8 + 2 * t4 + e];
          const int kb = ki & 0xffff;
          int x0, x1;
          if (oca) {
            x0 = kb - qb0;
            x1 = kb - qb1;
            if (x0 < 0) x0 += nb * nb;
            if (x1 < 0) x1 += nb * nb;
          } else {
            x0 = qb0 - kb;
            x1 = qb1 - kb;
          }
          float v0 = s[nt][e] + btab[x0], v1 = s[nt][2 + e] + btab[x1];
          if (masked) {
            const int rj = ki >> 16;
            if (rj != rid0) v0 += -100.0f;
            if (rj != rid1) v1 += -100.0f;
          }
          s[nt][e] = v0;
          s[nt][2 + e] = v1;
          c0 = fmaxf(c0, v0);
          c1 = fmaxf(c1, v1);
        }
      }
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1));
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
      const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);
      const float f0 = ex2_fast((m0 - n0) * LOG2E), f1 = ex2_fast((m1 - n1) * LOG2E);  // exp2(-inf) = 0 on the first chunk
      m0 = n0;
      m1 = n1;
      float r0 = 0.0f, r1 = 0.0f;
      const float nb0 = n0 * LOG2E, nb1 = n1 * LOG2E;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p0 = ex2_fast(fmaf(s[nt][e], LOG2E, -nb0)), p1 = ex2_fast(fmaf(s[nt][2 + e], LOG2E, -nb1));
          s[nt][e] = p0;
          s[nt][2 + e] = p1;
          r0 += p0;
          r1 += p1;
        }
      }
      r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
      r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
      r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
      r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
      l0 = l0 * f0 + r0;
      l1 = l1 * f1 + r1;
#pragma unroll
      for (int nt = 0; nt < DP / 8; ++nt) {
        o[nt][0] *= f0;
        o[nt][1] *= f0;
        o[nt][2] *= f1;
        o[nt][3] *= f1;
      }
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
          ldsm_x4_t(sV + (kc + kk * 16 + (mat & 1) * 8 + (lane & 7)) * RB + np * 32 + (mat >> 1) * 16, b0, b1, b2, b3);
          mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
          mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __syncwarp();  // this warp's Q fragments were read before the key loop; its 16 rows of tQ become output staging
    const int r0l = warp * 16 + g, r1l = r0l + 8;
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
      *reinterpret_cast<uint32_t*>(tQ + r0l * RB + (nt * 8 + 2 * t4) * 2) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
      *reinterpret_cast<uint32_t*>(tQ + r1l * RB + (nt * 8 + 2 * t4) * 2) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
    __syncthreads();
    for (int e = tid; e < 64 * CH; e += 128) {
      const int r = e / CH, c = e % CH;
      *reinterpret_cast<uint4*>(out + (size_t)qpix[qb + r] * a.ld_o + h * DP + c * 8) = *reinterpret_cast<const uint4*>(tQ + r * RB + c * 16);
    }
  }
}

int launch_attn_flash(const AttnArgs& a, int oca, cudaStream_t s) {
  const int kws = oca ? a.kws : a.ws;
  const int Nq = a.ws * a.ws, Nk = kws * kws, nb = a.ws + kws - 1;
  SSR_CHECK(a.DP == 16 || a.DP == 32, SSR_E_INVALID, "attn_flash: padded head dim %d not in {16,32}", a.DP);
  SSR_CHECK(Nq % 64 == 0 && Nk % 64 == 0, SSR_E_INVALID, "attn_flash: windows %d / %d need multiples of 64 tokens", a.ws, kws);
  SSR_CHECK(a.H % a.ws == 0 && a.W % a.ws == 0, SSR_E_INVALID, "attn_flash: %dx%d not a multiple of ws=%d", a.H, a.W, a.ws);
  SSR_CHECK(a.ld_qkv % 8 == 0 && a.QP % 8 == 0 && a.ld_o % 8 == 0 && nb * nb < 65536, SSR_E_INVALID, "attn_flash: unaligned leading dims");
  const size_t RB = (size_t)a.DP * 2 + 16;
  const size_t smem = 2 * (size_t)Nk * RB + 64 * RB + (size_t)nb * nb * 4 + (size_t)(2 * Nk + 2 * Nq) * 4;
  SSR_CHECK(smem <= 200 * 1024, SSR_E_INVALID, "attn_flash: %zu B of shared memory", smem);
  const int nwin = a.B * (a.H / a.ws) * (a.W / a.ws);
  ProfScope prof(oca ? "attn_oca_mma" : "attn_win_mma", 4.0 * nwin * Nq * (double)Nk * a.d * a.heads,
                 (double)nwin * (2.0 * Nq + 2.0 * Nk) * a.heads * a.d * 2, s);
  static size_t attr16 = 0, attr32 = 0;
  if (a.DP == 32) {
    if (smem > attr32) {
      SSR_CUDA(cudaFuncSetAttribute(attn_flash_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr32 = smem;
    }
    attn_flash_kernel<32><<<dim3(nwin, a.heads), 128, smem, s>>>(a, oca);
  } else {
    if (smem > attr16) {
      SSR_CUDA(cudaFuncSetAttribute(attn_flash_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr16 = smem;
    }
    attn_flash_kernel<16><<<dim3(nwin, a.heads), 128, smem, s>>>(a, oca);
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
