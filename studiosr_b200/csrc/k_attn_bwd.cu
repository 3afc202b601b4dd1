// Backward of the (shifted-)window attention core (swinir.py:83-102 with the roll / window_partition / window_reverse
// of swinir.py:154-168 as addressing), bf16 mma.sync tiles, for 8x8 windows.  It is the adjoint torch autograd builds
// for the reference (trainer.py:104):
//   P = softmax(q k^T + B + mask);  dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(P o dP));  dQ = dS K;  dK = dS^T Q
//   dB[h][i][j] = sum over windows of dS   (scattered into the [225][heads] table by bias_table_grad_kernel)
// One CTA = one head; it walks a strided list of windows so the bias gradient accumulates in shared memory and
// leaves through one atomicAdd per entry.  4 warps x 16 query rows; P and dS make one round trip through shared
// memory (bf16) so that the products reducing over QUERY rows (dK, dV) read them transposed with ldmatrix.trans.
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

template <int DP>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const AttnBwdArgs a) {
  constexpr int RB = DP * 2 + 16;  // bytes per row of the q / k / v / dO tiles (+16: ldmatrix rows hit different banks)
  constexpr int PB = 64 * 2 + 16;  // bytes per row of the P / dS tiles
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* tq = smem;             // [64][RB] q, later dq
  uint8_t* tk = tq + 64 * RB;     // k, later dk
  uint8_t* tv = tk + 64 * RB;     // v, later dv
  uint8_t* tdo = tv + 64 * RB;    // dO
  uint8_t* tP = tdo + 64 * RB;    // [64][PB] bf16 P
  uint8_t* tdS = tP + 64 * PB;    // [64][PB] bf16 dS
  float* dB = reinterpret_cast<float*>(tdS + 64 * PB);  // [64][64] fp32 accumulated over this CTA's windows
  float* btab = dB + 64 * 64;                           // [225]
  int* pix = reinterpret_cast<int*>(btab + 232);        // [64]
  int* rid = pix + 64;                                  // [64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y;
  const int nwx = a.W / 8, nwy = a.H / 8, nwin = a.B * nwx * nwy;
  for (int e = tid; e < 64 * 64; e += 128) dB[e] = 0.0f;
  for (int e = tid; e < 225; e += 128) btab[e] = __ldg(a.bias + h * 225 + e);
  const uint32_t sq = (uint32_t)__cvta_generic_to_shared(tq), sk = (uint32_t)__cvta_generic_to_shared(tk);
  const uint32_t sv = (uint32_t)__cvta_generic_to_shared(tv), sdo = (uint32_t)__cvta_generic_to_shared(tdo);
  const uint32_t sP = (uint32_t)__cvta_generic_to_shared(tP), sdS = (uint32_t)__cvta_generic_to_shared(tdS);
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv);
  const __nv_bfloat16* dO = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(a.dqkv);
  const int g = lane >> 2, t = lane & 3;
  const int i0 = warp * 16 + g, i1 = i0 + 8;
  const int yi0 = i0 >> 3, xi0 = i0 & 7, yi1 = i1 >> 3, xi1 = i1 & 7;
  const bool masked = a.shift > 0;
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr int CH = DP / 8;  // 16-byte chunks per tile row

  for (int wdw = blockIdx.x; wdw < nwin; wdw += gridDim.x) {
    __syncthreads();  // previous window's scatter has finished reading the tiles
    if (tid < 64) {
      int w = wdw;
      const int wx = w % nwx;
      w /= nwx;
      const int wy = w % nwy, b = w / nwy;
      const int sy = wy * 8 + tid / 8, sx = wx * 8 + tid % 8;
      const int yy = (sy + a.shift) % a.H, xx = (sx + a.shift) % a.W;
      pix[tid] = (b * a.H + yy) * a.W + xx;
      rid[tid] = a.shift > 0 ? 3 * shift_region(sy, a.H, 8, a.shift) + shift_region(sx, a.W, 8, a.shift) : 0;
    }
    __syncthreads();
    for (int e = tid; e < 64 * CH * 4; e += 128) {
      const int part = e / (64 * CH), r = (e / CH) % 64, c = e % CH;
      const uint32_t dst = (part == 0 ? sq : part == 1 ? sk : part == 2 ? sv : sdo) + r * RB + c * 16;
      const __nv_bfloat16* src = part < 3 ? qkv + (size_t)pix[r] * a.ld_qkv + part * a.QP + h * DP + c * 8
                                          : dO + (size_t)pix[r] * a.ld_do + h * DP + c * 8;
      cp_async16(dst, src);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- S = q k^T (+bias, mask), P = softmax(S) for this warp's 16 query rows ----
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
      uint32_t a0, a1, a2, a3, c0, c1, c2, c3;
      ldsm_x4(sq + (warp * 16 + (lane & 15)) * RB + ks * 32 + (lane >> 4) * 16, a0, a1, a2, a3);
      ldsm_x4(sdo + (warp * 16 + (lane & 15)) * RB + ks * 32 + (lane >> 4) * 16, c0, c1, c2, c3);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int mat = lane >> 3;
        const uint32_t off = (np * 16 + (mat >> 1) * 8 + (lane & 7)) * RB + ks * 32 + (mat & 1) * 16;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sk + off, b0, b1, b2, b3);
        mma_bf16(s[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
        ldsm_x4(sv + off, b0, b1, b2, b3);  // dP = dO v^T has the same operand pattern
        mma_bf16(dp[2 * np], c0, c1, c2, c3, b0, b1);
        mma_bf16(dp[2 * np + 1], c0, c1, c2, c3, b2, b3);
      }
    }
    const int rid0 = rid[i0], rid1 = rid[i1];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        const int yj = j >> 3, xj = j & 7;
        float v0 = s[nt][e] + btab[(yi0 - yj + 7) * 15 + (xi0 - xj + 7)];
        float v1 = s[nt][2 + e] + btab[(yi1 - yj + 7) * 15 + (xi1 - xj + 7)];
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
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = ex2_fast((s[nt][e] - m0) * LOG2E);
        s[nt][2 + e] = ex2_fast((s[nt][2 + e] - m1) * LOG2E);
        l0 += s[nt][e];
        l1 += s[nt][2 + e];
      }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    float d0 = 0.0f, d1 = 0.0f;  // delta_i = sum_j P dP
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] *= inv0;
        s[nt][2 + e] *= inv1;
        d0 = fmaf(s[nt][e], dp[nt][e], d0);
        d1 = fmaf(s[nt][2 + e], dp[nt][2 + e], d1);
      }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    // dS = P o (dP - delta); bias gradient; P and dS to shared memory (bf16) for the transposed products
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int j = nt * 8 + 2 * t;
      dp[nt][0] = s[nt][0] * (dp[nt][0] - d0);
      dp[nt][1] = s[nt][1] * (dp[nt][1] - d0);
      dp[nt][2] = s[nt][2] * (dp[nt][2] - d1);
      dp[nt][3] = s[nt][3] * (dp[nt][3] - d1);
      // column index XOR-swizzled by the row (i0 & 7 == i1 & 7 == g): the 8 rows a warp touches per access would all sit in
      // the same banks (row pitch 256 B) -- 8-way conflicts on 64 read-modify-writes per window
      float2* b0p = reinterpret_cast<float2*>(dB + i0 * 64 + (j ^ (g << 3)));
      float2* b1p = reinterpret_cast<float2*>(dB + i1 * 64 + (j ^ (g << 3)));
      float2 x = *b0p, y = *b1p;
      x.x += dp[nt][0]; x.y += dp[nt][1]; y.x += dp[nt][2]; y.y += dp[nt][3];
      *b0p = x;
      *b1p = y;
      *reinterpret_cast<uint32_t*>(tP + i0 * PB + j * 2) = pack_bf16x2(s[nt][0], s[nt][1]);
      *reinterpret_cast<uint32_t*>(tP + i1 * PB + j * 2) = pack_bf16x2(s[nt][2], s[nt][3]);
      *reinterpret_cast<uint32_t*>(tdS + i0 * PB + j * 2) = pack_bf16x2(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(tdS + i1 * PB + j * 2) = pack_bf16x2(dp[nt][2], dp[nt][3]);
    }
    // ---- dQ = dS K for this warp's rows (A = dS from registers, B = K read transposed) ----
    float dq[DP / 8][4];
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t a0 = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
      const uint32_t a1 = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      const uint32_t a3 = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < DP / 16; ++np) {
        const int mat = lane >> 3;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sk + (kk * 16 + (mat & 1) * 8 + (lane & 7)) * RB + np * 32 + (mat >> 1) * 16, b0, b1, b2, b3);
        mma_bf16(dq[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(dq[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();  // P / dS of all 64 query rows are in shared memory

    // ---- dK = dS^T Q, dV = P^T dO for key rows [16 warp, 16 warp + 16): A read transposed from the P / dS tiles ----
    float dk[DP / 8][4], dv[DP / 8][4];
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[nt][e] = dv[nt][e] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 query rows per step
      const int mat = lane >> 3;
      // A[m = key j][k = query i] = X[i][j]: matrix (rb = mat & 1, kb = mat >> 1) = rows i of the tile, transposed on load
      const uint32_t aoff = (kk * 16 + (mat >> 1) * 8 + (lane & 7)) * PB + (warp * 16 + (mat & 1) * 8) * 2;
      uint32_t a0, a1, a2, a3, c0, c1, c2, c3;
      ldsm_x4_t(sdS + aoff, a0, a1, a2, a3);
      ldsm_x4_t(sP + aoff, c0, c1, c2, c3);
#pragma unroll
      for (int np = 0; np < DP / 16; ++np) {
        const uint32_t boff = (kk * 16 + (mat & 1) * 8 + (lane & 7)) * RB + np * 32 + (mat >> 1) * 16;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sq + boff, b0, b1, b2, b3);
        mma_bf16(dk[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(dk[2 * np + 1], a0, a1, a2, a3, b2, b3);
        ldsm_x4_t(sdo + boff, b0, b1, b2, b3);
        mma_bf16(dv[2 * np], c0, c1, c2, c3, b0, b1);
        mma_bf16(dv[2 * np + 1], c0, c1, c2, c3, b2, b3);
      }
    }
    __syncthreads();  // every warp is done reading q / k / v / dO: the tiles become the output staging
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
      const int cb = (nt * 8 + 2 * t) * 2;
      *reinterpret_cast<uint32_t*>(tq + i0 * RB + cb) = pack_bf16x2(dq[nt][0], dq[nt][1]);
      *reinterpret_cast<uint32_t*>(tq + i1 * RB + cb) = pack_bf16x2(dq[nt][2], dq[nt][3]);
      *reinterpret_cast<uint32_t*>(tk + i0 * RB + cb) = pack_bf16x2(dk[nt][0], dk[nt][1]);
      *reinterpret_cast<uint32_t*>(tk + i1 * RB + cb) = pack_bf16x2(dk[nt][2], dk[nt][3]);
      *reinterpret_cast<uint32_t*>(tv + i0 * RB + cb) = pack_bf16x2(dv[nt][0], dv[nt][1]);
      *reinterpret_cast<uint32_t*>(tv + i1 * RB + cb) = pack_bf16x2(dv[nt][2], dv[nt][3]);
    }
    __syncthreads();
    for (int e = tid; e < 64 * CH * 3; e += 128) {
      const int part = e / (64 * CH), r = (e / CH) % 64, c = e % CH;
      const uint8_t* src = (part == 0 ? tq : part == 1 ? tk : tv) + r * RB + c * 16;
      *reinterpret_cast<uint4*>(dqkv + (size_t)pix[r] * a.ld_qkv + part * a.QP + h * DP + c * 8) = *reinterpret_cast<const uint4*>(src);
    }
  }
  __syncthreads();
  float* out = a.dB + (size_t)h * 64 * 64;
  for (int e = tid; e < 64 * 64; e += 128) atomicAdd(out + e, dB[(e & ~63) + ((e & 63) ^ (((e >> 6) & 7) << 3))]);
}

// dB [heads][64][64] -> d(relative_position_bias_table) [225][heads] (swinir.py:57-67, 92-95): entries (i, j) with the
// same relative offset share one table row
__global__ void bias_table_grad_kernel(const float* __restrict__ dB, float* dtable, int heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 225 * heads) return;
  const int h = idx % heads, rel = idx / heads;
  const int dy = rel / 15 - 7, dx = rel % 15 - 7;
  float acc = 0.0f;
  for (int yi = 0; yi < 8; ++yi) {
    const int yj = yi - dy;
    if (yj < 0 || yj > 7) continue;
    for (int xi = 0; xi < 8; ++xi) {
      const int xj = xi - dx;
      if (xj < 0 || xj > 7) continue;
      acc += dB[((size_t)h * 64 + yi * 8 + xi) * 64 + yj * 8 + xj];
    }
  }
  dtable[idx] = acc;
}

int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t s) {
  SSR_CHECK(a.DP == 16 || a.DP == 32, SSR_E_INVALID, "attn_bwd: padded head dim %d not in {16,32}", a.DP);
  SSR_CHECK(a.H % 8 == 0 && a.W % 8 == 0, SSR_E_INVALID, "attn_bwd: %dx%d not a multiple of 8", a.H, a.W);
  SSR_CHECK(a.ld_qkv % 8 == 0 && a.QP % 8 == 0 && a.ld_do % 8 == 0, SSR_E_INVALID, "attn_bwd: unaligned leading dims");
  const int nwin = a.B * (a.H / 8) * (a.W / 8);
  const size_t smem = (size_t)4 * 64 * (a.DP * 2 + 16) + 2 * 64 * (64 * 2 + 16) + 64 * 64 * 4 + 232 * 4 + 128 * 4;
  SSR_CUDA(cudaMemsetAsync(a.dB, 0, (size_t)a.heads * 64 * 64 * 4, s));
  int groups = (4 * 148 + a.heads - 1) / a.heads;
  if (groups > nwin) groups = nwin;
  ProfScope prof("attn_bwd_mma", 10.0 * nwin * 64 * 64 * a.d * a.heads, 7.0 * nwin * 64 * a.heads * a.d * 2, s);
  static bool attr16 = false, attr32 = false;
  if (a.DP == 32) {
    if (!attr32) {
      SSR_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr32 = true;
    }
    attn_bwd_kernel<32><<<dim3(groups, a.heads), 128, smem, s>>>(a);
  } else {
    if (!attr16) {
      SSR_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr16 = true;
    }
    attn_bwd_kernel<16><<<dim3(groups, a.heads), 128, smem, s>>>(a);
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  if (a.dtable) {
    bias_table_grad_kernel<<<(225 * a.heads + 127) / 128, 128, 0, s>>>(a.dB, a.dtable, a.heads);
    count_launch();
    SSR_CUDA(cudaGetLastError());
  }
  return SSR_OK;
}

}  // namespace ssr
