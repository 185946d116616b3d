// osc_b200.cu -- sm_100a kernels and the C-ABI (include/osc_b200.h) of the batched
// operational-space controller.
//
// Kernels
//   build_qp_kernel   K2: objective matrices H (dv block) and f from J, bias, targets.
//                     HBM-bound stream: every environment's J/bias/targets record is
//                     pulled into shared memory by 1-D TMA bulk copies (cp.async.bulk,
//                     mbarrier completion) through a 2-stage ring per warp; the products
//                     J'WJ and J'W(bias - t) run on the FP64 tensor cores (DMMA).
//                     Stand-alone form: osc_setup, the condensed mode, the three-kernel step.
//   init_state_kernel set_up_optimization(): cold iterates, rho0, first linear cost.
//   build_scale_kernel3
//                     K2 + K3a fused (the default first kernel of a control step): the warp
//                     lands M, J, bias, targets, builds H and f in shared memory (DMMA), sends
//                     them out with bulk stores and runs the Ruiz passes on registers.
//   scale_kernel3     K3a alone: OSQP's scale_data (Ruiz equilibration + cost scaling) and the
//                     update-path decision, one warp per environment, unscaled matrices in
//                     registers for all passes (osc::Core3::ruiz).
//   solve_kernel3     K3b: one warp per environment, persistent CTAs with a dynamic work
//                     counter and a per-warp landing stage (the next environment's TMA bulk
//                     copies overlap the solve of the current one): assembly, factorisation,
//                     ADMM with register-resident matrices (two exchanges per iteration),
//                     un-scaling, entirely on chip in FP64 (osc::Core3::step_prepare /
//                     step_solve).
//   condensed_kernel  K1 + K3 of the condensed fast mode (osc_condensed.cuh), reported separately.
//   kinematics_kernel M, C, J, Jdot qvel of a floating-base hinge tree (osc_kinematics.cuh).
//   order_kernel      longest-first hand-out order of solve_kernel3's work counter (scheduling).
//   reset_warm_kernel reset_optimization().
//   targets_pd_kernel, targets_walter_tumbling_kernel, contact_mask_kernel
//                     the step before the path for device-resident roll-outs: task-space
//                     targets and contact masks.
//   widen_rows_kernel FP32 -> FP64 rows of J (opt-in FP32 transport, osc_step_host_j32).
//   gather_push_kernel
//                     torques + statistics of a rank into every rank's slab (peer stores).
//   warp_selftest_kernel
//                     the device reading of the warp primitives of osc_warp.cuh (tests).
//   dfma_peak_kernel  FP64-FMA roofline denominator.
//
// No CPU fallback exists in this library.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#ifdef OSC_PHASE_CLOCKS  // developer build only (tools/phase_clocks.py)
__device__ unsigned long long g_phase_clocks[64];
__host__ __device__ __forceinline__ void osc_tick(int k, int lane0) {
#if defined(__CUDA_ARCH__)
  if (lane0 == 0) atomicAdd(&g_phase_clocks[k], (unsigned long long)clock64());
#else
  (void)k;
  (void)lane0;
#endif
}
#define OSC_TICK(k) osc_tick(k, lane0)
#define OSC_TICKL(k) osc_tick(k, lane)
#else
#define OSC_TICKL(k) ((void)0)
#endif

#include "osc_params.h"
#include "osc_condensed.cuh"
#include "osc_kinematics.cuh"

namespace osc {

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA unit, UBLKCP in SASS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// shared -> global bulk copy (bulk-group completion), and its two waits: `read` = the source
// may be overwritten, `all` = the data is in global memory
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------
// K2: objective build.  One warp per environment; every warp owns a 2-stage ring of TMA
// bulk copies (J, bias, targets of one environment per stage).
//   H = 2 J'WJ + 2 w_reg I   is a small GEMM  sum_k (w_k J_k)' J_k  with K = 6 ns rows, and
//   f = 2 J'W(bias - t)      is the same product against one more column r = bias - t,
// so both run on the FP64 tensor cores: mma.sync m8n8k4 (DMMA), 8x8 output tiles of the
// lower triangle of H.  nv is not a multiple of 8, so the last 8-column block of J has
// zero padding: r is placed in padding column nv, which makes row (nv mod 8) of the last
// tile row equal to f/2 -- f costs no extra MMA.
// Per 4 rows of J a lane loads one element per 8-column block (A and B fragments of a
// tile are the same J entries), scales the A copy by the row weight and issues the MMAs.
// ---------------------------------------------------------------------------
template <class D>
struct BuildStage {
  double J[D::S * D::NV];
  double bias[D::S];
  double targets[D::S];
};
constexpr int kBuildStages = 2;
constexpr int kBuildWarps = 4;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// The k-loop of the objective build: acc = lower-triangle 8x8 tiles of sum_k (w_k Jx_k)' Jx_k
// with Jx = [J | r | 0] (r = bias - t rides in padding column nv of the last 8-column block).
template <class D>
__device__ __forceinline__ void build_accumulate(
    const double* __restrict__ Js, const double* __restrict__ rs, const double* __restrict__ w_row,
    int lane, double (&acc)[((D::NV + 7) / 8) * ((D::NV + 7) / 8 + 1) / 2][2]) {
  constexpr int NV = D::NV, S = D::S;
  constexpr int NB8 = (NV + 7) / 8;
  constexpr int NTILE = NB8 * (NB8 + 1) / 2;
  constexpr int KSTEPS = (S + 3) / 4;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int q = 0; q < NTILE; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 2
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const int k = 4 * ks + t;            // this lane's row of J inside the k-step
    const int kc = k < S ? k : S - 1;    // clamp (weight is 0 beyond S)
    const double wk = w_row[k];
    double jb[NB8], ja[NB8];
#pragma unroll
    for (int q = 0; q < NB8; ++q) {
      const int col = 8 * q + g;
      if (q < NB8 - 1) {
        jb[q] = Js[kc * NV + col];
      } else {
        // last block: columns < nv from J, column nv carries r, the rest is zero
        const double* src = col < NV ? &Js[kc * NV + col] : &rs[kc];
        jb[q] = col <= NV ? *src : 0.0;
      }
      ja[q] = wk * jb[q];
    }
    int tile = 0;
#pragma unroll
    for (int mi = 0; mi < NB8; ++mi)
#pragma unroll
      for (int ni = 0; ni <= mi; ++ni, ++tile) dmma_m8n8k4(acc[tile][0], acc[tile][1], ja[mi], jb[ni]);
  }
}

template <class D>
__global__ void __launch_bounds__(kBuildWarps * 32)
build_qp_kernel(const __grid_constant__ Params p, const double* __restrict__ J,
                const double* __restrict__ bias, const double* __restrict__ targets,
                double* __restrict__ Hdv, double* __restrict__ fdv, int n_envs) {
  constexpr int NV = D::NV, S = D::S, NS = D::NS;
  constexpr int NB8 = (NV + 7) / 8;              // 8-column blocks of J
  constexpr int NTILE = NB8 * (NB8 + 1) / 2;     // lower-triangle 8x8 tiles of H
  constexpr int KSTEPS = (S + 3) / 4;
  static_assert(NV % 8 != 0, "f rides in the zero padding of the last 8-column block");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BuildStage<D>* stages = reinterpret_cast<BuildStage<D>*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kBuildWarps * kBuildStages * sizeof(BuildStage<D>));
  double* w_row = reinterpret_cast<double*>(bars + kBuildWarps * kBuildStages);  // 4*KSTEPS, zero padded
  int* t_idx = reinterpret_cast<int*>(w_row + 4 * KSTEPS);  // index into targets of row k

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;  // MMA fragment coordinates
  BuildStage<D>* my = stages + warp * kBuildStages;
  uint64_t* mybar = bars + warp * kBuildStages;
  constexpr uint32_t kBytes = sizeof(BuildStage<D>);
  auto issue = [&](int stage, int env) {
    BuildStage<D>& st = my[stage];
    // the stage was read and (bias -> r, in place) written through the generic proxy: order
    // those accesses before the async-proxy writes of the bulk copies
    fence_proxy_async();
    mbar_expect_tx(&mybar[stage], kBytes);
    bulk_g2s(st.J, J + (size_t)env * S * NV, sizeof(st.J), &mybar[stage]);
    bulk_g2s(st.bias, bias + (size_t)env * S, sizeof(st.bias), &mybar[stage]);
    bulk_g2s(st.targets, targets + (size_t)env * S, sizeof(st.targets), &mybar[stage]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBuildWarps * kBuildStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int k = threadIdx.x; k < 4 * KSTEPS; k += blockDim.x) {
    w_row[k] = k < S ? p.w_row[k] : 0.0;
    // row k of ddx: site (k mod 3NS)/3, translational rows read targets cols 0-2,
    // rotational rows cols 3-5 (autogen.py:163,173-177)
    const int kc = k < S ? k : S - 1;
    const int kr = (kc < 3 * NS) ? kc : kc - 3 * NS;
    t_idx[k] = (kr / 3) * 6 + (kr % 3) + ((kc < 3 * NS) ? 0 : 3);
  }
  __syncthreads();
  const int stride = gridDim.x * kBuildWarps;
  const int env0 = blockIdx.x * kBuildWarps + warp;
  if (lane == 0) {
    for (int s = 0; s < kBuildStages; ++s)
      if (env0 + s * stride < n_envs) issue(s, env0 + s * stride);
  }
  const double two_wreg = 2.0 * p.w_reg;
  int it = 0;
  for (int env = env0; env < n_envs; env += stride, ++it) {
    const int stage = it % kBuildStages;
    mbar_wait(&mybar[stage], (it / kBuildStages) & 1);
    const BuildStage<D>& st = my[stage];
    // r = bias - t, in place over the staged bias (t re-ordered [translational ; rotational])
    {
      double* b = const_cast<double*>(st.bias);
      for (int k = lane; k < S; k += 32) b[k] = b[k] - st.targets[t_idx[k]];
      __syncwarp();
    }
    double acc[NTILE][2];
    build_accumulate<D>(st.J, st.bias, w_row, lane, acc);
    // C fragment: lane holds C[g][2t], C[g][2t+1] of every tile
    double* H = Hdv + (size_t)env * NV * NV;
    int tile = 0;
#pragma unroll
    for (int mi = 0; mi < NB8; ++mi) {
      const int row = 8 * mi + g;
#pragma unroll
      for (int ni = 0; ni <= mi; ++ni, ++tile) {
        const int col = 8 * ni + 2 * t;
        double v0 = 2.0 * acc[tile][0], v1 = 2.0 * acc[tile][1];
        if (row < NV && col < NV) {
          if (mi == ni) {
            // keep the symmetric matrix exactly symmetric: the lower-triangle value is used
            // for both (i,j) and (j,i), like the reference's mirrored Hessian
            if (col <= row) {
              if (col == row) v0 += two_wreg;
              H[row * NV + col] = v0;
              if (col != row) H[col * NV + row] = v0;
            }
            if (col + 1 <= row) {
              if (col + 1 == row) v1 += two_wreg;
              H[row * NV + col + 1] = v1;
              if (col + 1 != row) H[(col + 1) * NV + row] = v1;
            }
          } else {
            *reinterpret_cast<double2*>(&H[row * NV + col]) = make_double2(v0, v1);
            H[col * NV + row] = v0;
            H[(col + 1) * NV + row] = v1;
          }
        }
      }
    }
    // f/2 = row (nv mod 8) of the last tile row: sum_k (w_k r_k) J[k][col]
    if (g == NV % 8) {
#pragma unroll
      for (int ni = 0; ni < NB8; ++ni) {
        const int q = (NB8 - 1) * NB8 / 2 + ni;
        const int col = 8 * ni + 2 * t;
        if (col < NV) fdv[(size_t)env * NV + col] = 2.0 * acc[q][0];
        if (col + 1 < NV) fdv[(size_t)env * NV + col + 1] = 2.0 * acc[q][1];
      }
    }
    __syncwarp();  // the whole warp is done reading this stage
    if (lane == 0) {
      const int next = env + kBuildStages * stride;
      if (next < n_envs) issue(stage, next);
    }
  }
}

// ---------------------------------------------------------------------------
// set_up_optimization(): Init state
// ---------------------------------------------------------------------------
// One warp per environment: zero iterates, previous linear cost = f, rho0, flag, and the
// sparsity signature of the data the "workspace" is initialised with.
template <class D>
__global__ void init_state_kernel(double* __restrict__ state, const double* __restrict__ fdv,
                                  const double* __restrict__ Hdv, const double* __restrict__ M,
                                  const double* __restrict__ J, double rho0, int n_envs) {
  constexpr int NV = D::NV, QOFF = D::N + 2 * D::M;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; env < n_envs; env += warps) {
    double* st = state + (size_t)env * D::STATE;
    for (int k = lane; k < QOFF; k += 32) st[k] = 0.0;
    for (int k = lane; k < NV; k += 32) st[QOFF + k] = fdv[(size_t)env * NV + k];
    if (lane == 0) {
      st[QOFF + NV] = rho0;
      st[QOFF + NV + 1] = 1.0;
    }
    const double* H = Hdv + (size_t)env * NV * NV;
    const double* Mm = M + (size_t)env * NV * NV;
    const double* Jc = J + ((size_t)env * D::S + D::JC0) * NV;
    auto bit = [&](int b) -> bool {
      if (b < NV * NV) return H[b] != 0.0;
      b -= NV * NV;
      if (b < NV * NV) return Mm[b] != 0.0;
      b -= NV * NV;
      if (b < NV * D::NZ) return Jc[b] != 0.0;
      return false;
    };
    for (int q = 0; q < D::SIG; ++q) {
      const unsigned lo = __ballot_sync(0xffffffffu, bit(64 * q + lane));
      const unsigned hi = __ballot_sync(0xffffffffu, bit(64 * q + 32 + lane));
      if (lane == 0) st[D::SIG0 + q] = __longlong_as_double(((long long)hi << 32) | (long long)lo);
    }
  }
}

template <class D>
__global__ void reset_warm_kernel(double* __restrict__ state, int n_envs) {
  constexpr int ST = D::STATE, QOFF = D::N + 2 * D::M;
  const size_t total = (size_t)n_envs * QOFF;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int env = (int)(i / QOFF), k = (int)(i - (size_t)env * QOFF);
    state[(size_t)env * ST + k] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// K3: solve
// ---------------------------------------------------------------------------
struct SolveArgs {
  const double *M, *C, *J, *mask, *Hdv, *fdv;
  double* state;
  double *torque, *sol_x, *sol_y, *pri_res, *dua_res, *rho;
  int *iters, *status;
  const double* scal;  // scaling records of scale_kernel3
  // work counter: never reset -- a launch over n environments with W warps draws exactly
  // n + 2 W tickets (every warp draws two past its work: tickets are drawn ahead of their
  // use), so the host knows where the next launch's tickets start (`base`) without a memset
  // between the kernels
  unsigned* counter;
  unsigned base;
  // ticket -> environment (a permutation of [0, n_envs), longest expected solves first), or
  // NULL for the identity
  const int* order;
  int* reinits;
  int n_envs;
};

// K3a for robots with 2 nv <= 32: OSQP's scale_data (10 Ruiz passes) + the update-path decision
// of update_optimization (:565-584), osc::Core3::ruiz.  Its own kernel because it needs a
// third of the registers of the solve (more warps per SM hide its dependent max / rsqrt
// chains): one warp per environment, persistent CTAs, dynamic work counter, the next
// environment's bulk copies in flight while the passes run on registers.
struct ScaleArgs {
  const double *M, *J, *Hdv, *fdv;
  const double *bias, *targets;  // fused build only
  double *Hdv_out, *fdv_out;     // fused build only: H, f for the solve kernel
  double* state;  // reads previous f / flag / signature, updates the signature in place
  double* scal;   // out: D, E, c, path flag per environment
  unsigned* counter;  // see SolveArgs
  unsigned base;
  int n_envs;
};

template <class D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
scale_kernel3(const __grid_constant__ Params p, const ScaleArgs a) {
  using RWS = RuizWorkspace<D>;
  using C3 = Core3<D>;
  constexpr int NV = D::NV, TAIL0 = D::N + 2 * D::M;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RWS* wsb = reinterpret_cast<RWS*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + WARPS * sizeof(RWS));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RWS& rw = wsb[warp];
  uint64_t* bar = &bars[warp];
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  constexpr uint32_t kBytes = sizeof(typename RWS::Stage);
  static_assert(sizeof(typename RWS::Stage) ==
                    sizeof(double) * (2 * NV * NV + D::NZ * NV + RWS::TAIL + NV),
                "the landing stage is exactly the five bulk copies");
  // tickets are drawn one environment ahead of their use (`ticket`: lane 0 only), so that the
  // round trip of the atomic never stalls the warp; every warp draws two tickets past its work
  auto draw = [&]() -> unsigned {  // only the atomic: its result is first touched in fetch()
    unsigned raw = 0;
    if (lane == 0) raw = atomicAdd(a.counter, 1u);
    return raw;
  };
  auto fetch = [&](unsigned raw) -> int {
    const int ticket = (int)(raw - a.base);
    if (lane == 0 && ticket < a.n_envs) {
      const int env = ticket;
      fence_proxy_async();
      mbar_expect_tx(bar, kBytes);
      bulk_g2s(rw.in.M, a.M + (size_t)env * NV * NV, sizeof(rw.in.M), bar);
      bulk_g2s(rw.in.H, a.Hdv + (size_t)env * NV * NV, sizeof(rw.in.H), bar);
      bulk_g2s(rw.in.Jc, a.J + ((size_t)env * D::S + D::JC0) * NV, sizeof(rw.in.Jc), bar);
      bulk_g2s(rw.in.tail, a.state + (size_t)env * D::STATE + TAIL0, sizeof(rw.in.tail), bar);
      bulk_g2s(rw.in.fv, a.fdv + (size_t)env * NV, sizeof(rw.in.fv), bar);
    }
    return __shfl_sync(0xffffffffu, ticket, 0);
  };
  uint32_t parity = 0;
  int env = fetch(draw());
  unsigned ticket = draw();
  while (env < a.n_envs) {
    mbar_wait(bar, parity);
    parity ^= 1;
    int next = a.n_envs;
    C3::ruiz(rw, p, lane, a.scal + (size_t)env * C3::SCAL,
             a.state + (size_t)env * D::STATE + D::SIG0, [&]() {
               next = fetch(ticket);
               ticket = draw();
             });
    env = next;
  }
}

// K2 + K3a fused: the objective build (DMMA, as in build_qp_kernel) runs in the warp that
// equilibrates the environment, on the task Jacobian it lands anyway for the contact rows.
// H and f never make a round trip through HBM before the Ruiz passes (they leave through two
// bulk stores for the solve kernel), one launch and one kernel tail less per control step,
// and the HBM-bound read of J hides under the latency-bound passes of the other warps.
template <class D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
build_scale_kernel3(const __grid_constant__ Params p, const ScaleArgs a) {
  using RWS = BuildRuizWorkspace<D>;
  using C3 = Core3<D>;
  constexpr int NV = D::NV, S = D::S, NS = D::NS, TAIL0 = D::N + 2 * D::M;
  constexpr int NB8 = (NV + 7) / 8, NTILE = NB8 * (NB8 + 1) / 2, KSTEPS = (S + 3) / 4;
  static_assert(NV % 8 != 0, "f rides in the zero padding of the last 8-column block");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RWS* wsb = reinterpret_cast<RWS*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + WARPS * sizeof(RWS));
  double* w_row = reinterpret_cast<double*>(bars + WARPS);  // 4 * KSTEPS, zero padded
  int* t_idx = reinterpret_cast<int*>(w_row + 4 * KSTEPS);  // index into targets of row k
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RWS& rw = wsb[warp];
  uint64_t* bar = &bars[warp];
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (int k = threadIdx.x; k < 4 * KSTEPS; k += blockDim.x) {
    w_row[k] = k < S ? p.w_row[k] : 0.0;
    const int kc = k < S ? k : S - 1;
    const int kr = (kc < 3 * NS) ? kc : kc - 3 * NS;
    t_idx[k] = (kr / 3) * 6 + (kr % 3) + ((kc < 3 * NS) ? 0 : 3);  // autogen.py:163,173-177
  }
  __syncthreads();
  constexpr uint32_t kBytes = sizeof(typename RWS::Stage);
  static_assert(sizeof(typename RWS::Stage) ==
                    sizeof(double) * (NV * NV + S * NV + 2 * S + RWS::TAIL),
                "the landing stage is exactly the five bulk copies");
  // (tickets one environment ahead of their use, see scale_kernel3)
  auto draw = [&]() -> unsigned {  // only the atomic: its result is first touched in fetch()
    unsigned raw = 0;
    if (lane == 0) raw = atomicAdd(a.counter, 1u);
    return raw;
  };
  auto fetch = [&](unsigned raw) -> int {
    const int ticket = (int)(raw - a.base);
    if (lane == 0 && ticket < a.n_envs) {
      const int env = ticket;
      fence_proxy_async();
      mbar_expect_tx(bar, kBytes);
      bulk_g2s(rw.in.M, a.M + (size_t)env * NV * NV, sizeof(rw.in.M), bar);
      bulk_g2s(rw.in.J, a.J + (size_t)env * S * NV, sizeof(rw.in.J), bar);
      bulk_g2s(rw.in.bias, a.bias + (size_t)env * S, sizeof(rw.in.bias), bar);
      bulk_g2s(rw.in.targets, a.targets + (size_t)env * S, sizeof(rw.in.targets), bar);
      bulk_g2s(rw.in.tail, a.state + (size_t)env * D::STATE + TAIL0, sizeof(rw.in.tail), bar);
    }
    return __shfl_sync(0xffffffffu, ticket, 0);
  };
  const double two_wreg = 2.0 * p.w_reg;
  const int g = lane >> 2, t = lane & 3;
  uint32_t parity = 0;
  int env = fetch(draw());
  unsigned ticket = draw();
  while (env < a.n_envs) {
    OSC_TICKL(40);
    mbar_wait(bar, parity);
    parity ^= 1;
    OSC_TICKL(41);
    // ---- objective build: r = bias - t in place, then H, f in shared memory
    {
      constexpr int RR = (S + 31) / 32;
      double rv[RR];  // all loads, then all stores (the stores alias the loads for the compiler)
#pragma unroll
      for (int j = 0; j < RR; ++j) {
        const int k = lane + 32 * j;
        rv[j] = k < S ? rw.in.bias[k] - rw.in.targets[t_idx[k]] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < RR; ++j) {
        const int k = lane + 32 * j;
        if (k < S) rw.in.bias[k] = rv[j];
      }
    }
    if (lane == 0) bulk_wait_read();  // the previous environment's H, f have left
    __syncwarp();
    {
      double acc[NTILE][2];
      OSC_TICKL(42);
      build_accumulate<D>(rw.in.J, rw.in.bias, w_row, lane, acc);
      OSC_TICKL(43);
      int tile = 0;
#pragma unroll
      for (int mi = 0; mi < NB8; ++mi) {
        const int row = 8 * mi + g;
#pragma unroll
        for (int ni = 0; ni <= mi; ++ni, ++tile) {
          const int col = 8 * ni + 2 * t;
          double v0 = 2.0 * acc[tile][0], v1 = 2.0 * acc[tile][1];
          if (row < NV && col < NV) {
            if (mi == ni) {
              // the lower-triangle value is used for (i,j) and (j,i): exactly symmetric H
              if (col <= row) {
                if (col == row) v0 += two_wreg;
                rw.H[row * NV + col] = v0;
                if (col != row) rw.H[col * NV + row] = v0;
              }
              if (col + 1 <= row) {
                if (col + 1 == row) v1 += two_wreg;
                rw.H[row * NV + col + 1] = v1;
                if (col + 1 != row) rw.H[(col + 1) * NV + row] = v1;
              }
            } else {
              *reinterpret_cast<double2*>(&rw.H[row * NV + col]) = make_double2(v0, v1);
              rw.H[col * NV + row] = v0;
              rw.H[(col + 1) * NV + row] = v1;
            }
          }
        }
      }
      if (g == NV % 8) {  // f/2 = row (nv mod 8) of the last tile row
#pragma unroll
        for (int ni = 0; ni < NB8; ++ni) {
          const int q = (NB8 - 1) * NB8 / 2 + ni;
          const int col = 8 * ni + 2 * t;
          if (col < NV) rw.fv[col] = 2.0 * acc[q][0];
          if (col + 1 < NV) rw.fv[col + 1] = 2.0 * acc[q][1];
        }
      }
    }
    fence_proxy_async();  // H, f (generic-proxy stores) before the bulk stores read them
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(a.Hdv_out + (size_t)env * NV * NV, rw.H, sizeof(rw.H));
      bulk_s2g(a.fdv_out + (size_t)env * NV, rw.fv, sizeof(rw.fv));
      bulk_commit();
    }
    OSC_TICKL(44);
    int next = a.n_envs;
    C3::ruiz(rw, p, lane, a.scal + (size_t)env * C3::SCAL,
             a.state + (size_t)env * D::STATE + D::SIG0, [&]() {
               next = fetch(ticket);
               ticket = draw();
             });
    OSC_TICKL(45);
    env = next;
  }
  if (lane == 0) bulk_wait_all();
}

// K3 for robots with 2 nv <= 32: osc::Core3 (register-resident iteration matrices, two lanes
// per dynamics row).  One warp per environment, persistent CTAs, dynamic work counter.  Every
// warp owns a landing stage for one environment's input record; as soon as step_prepare() has
// consumed it the warp draws its next environment and lands it there with TMA bulk copies
// while step_solve() factorises and iterates on the current one.
constexpr int max_regs_for(int warps) {  // registers per thread that let `warps` warps share an SM
  const int r = (65536 / (warps * 32)) / 8 * 8;
  return r > 255 ? 255 : r;
}
template <class D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(max_regs_for(WARPS))
solve_kernel3(const __grid_constant__ Params p, const SolveArgs a) {
  using WS = Workspace3<D>;
  using C3 = Core3<D>;
  constexpr int NV = D::NV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WS* wsb = reinterpret_cast<WS*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + WARPS * sizeof(WS));
  // (opaque: under register pressure the compiler otherwise re-derives the workspace address
  // from %tid inside the iteration loop -- four S2R per iteration on the Go2 shape)
  // (one lane per row only: the two-lane shape keeps the address in a register by itself and
  // loses seven instructions per iteration to moves with the opaque ids)
  constexpr bool kOpaqueIds = 2 * D::NV > 32;
  const int warp = kOpaqueIds ? osc_opaque((int)(threadIdx.x >> 5)) : (int)(threadIdx.x >> 5);
  const int lane = kOpaqueIds ? osc_opaque((int)(threadIdx.x & 31)) : (int)(threadIdx.x & 31);
  WS& w = wsb[warp];
  uint64_t* bar = &bars[warp];
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  constexpr uint32_t kBytes = sizeof(typename WS::Stage);
  static_assert(sizeof(typename WS::Stage) ==
                    sizeof(double) * (2 * NV * NV + D::NZ * NV + D::STATE + 2 * NV + D::NC +
                                      C3::SCAL),
                "the landing stage is exactly the eight bulk copies");
  // lane 0 draws the next environment from the work counter (`draw`: issued early, the
  // round trip of the atomic overlaps step_prepare) and later starts landing it (`land`:
  // eight bulk copies into the stage; returns the index to all lanes)
  // Work tickets (lane 0).  `draw` only issues the atomic: nothing touches its result until
  // `resolve` turns it into an environment index (ticket - base, then the hand-out order), so
  // the round trip never stalls the warp.  A ticket is drawn just before the output stores of
  // one environment, resolved at the top of the next (the dependent load of the order then
  // has the whole assembly to arrive) and landed after that environment's assembly.  Every
  // warp draws two tickets past its work.
  auto draw = [&]() -> unsigned {
    unsigned raw = 0;
    if (lane == 0) raw = atomicAdd(a.counter, 1u);
    return raw;
  };
  auto resolve = [&](unsigned raw) -> int {
    const int t = (int)(raw - a.base);
    return (lane == 0 && a.order && t < a.n_envs) ? a.order[t] : t;
  };
  auto land = [&](int env) -> int {
    if (lane == 0 && env < a.n_envs) {
      fence_proxy_async();  // the stage's generic-proxy reads are ordered before the copies
      mbar_expect_tx(bar, kBytes);
      bulk_g2s(w.in.M, a.M + (size_t)env * NV * NV, sizeof(w.in.M), bar);
      bulk_g2s(w.in.H, a.Hdv + (size_t)env * NV * NV, sizeof(w.in.H), bar);
      bulk_g2s(w.in.Jc, a.J + ((size_t)env * D::S + D::JC0) * NV, sizeof(w.in.Jc), bar);
      bulk_g2s(w.in.land, a.state + (size_t)env * D::STATE, sizeof(w.in.land), bar);
      bulk_g2s(w.in.Cv, a.C + (size_t)env * NV, sizeof(w.in.Cv), bar);
      bulk_g2s(w.in.fv, a.fdv + (size_t)env * NV, sizeof(w.in.fv), bar);
      bulk_g2s(w.in.maskv, a.mask + (size_t)env * D::NC, sizeof(w.in.maskv), bar);
      bulk_g2s(w.in.scal, a.scal + (size_t)env * C3::SCAL, sizeof(w.in.scal), bar);
    }
    return __shfl_sync(0xffffffffu, env, 0);
  };
  uint32_t parity = 0;
#ifdef OSC_PHASE_CLOCKS  // SM clock rate while this kernel runs: cycles [29] per nanosecond [30]
  unsigned long long pc_c0 = 0, pc_t0 = 0;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    pc_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pc_t0));
  }
#endif
  int env = land(resolve(draw()));
  unsigned ticket = draw();
  while (env < a.n_envs) {
    const int drawn = resolve(ticket);
    mbar_wait(bar, parity);
    parity ^= 1;
    const int lane0 = lane;
    (void)lane0;
    OSC_TICK(0);
    typename C3::Regs L;
    double* sx = a.sol_x + (size_t)env * D::N;
    double* sy = a.sol_y + (size_t)env * D::M;
    double* so = a.state + (size_t)env * D::STATE;
    const typename C3::Prepared pr = C3::step_prepare(w, p, L, lane, sx, sy, so);
    __syncwarp();
    OSC_TICK(1);
    const int next = land(drawn);
    OSC_TICK(2);
    const Result r = C3::step_solve(w, p, L, lane, pr, sx, sy, a.torque + (size_t)env * D::NU,
                                    so, [&]() { ticket = draw(); });
    if (lane == 0) {
      a.iters[env] = r.iter;
      a.status[env] = r.status;
      a.pri_res[env] = r.pri_res;
      a.dua_res[env] = r.dua_res;
      a.rho[env] = r.rho;
      if (r.reinit) atomicAdd(a.reinits, 1);
    }
    __syncwarp();
    OSC_TICK(8);
    env = next;
  }
#ifdef OSC_PHASE_CLOCKS
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicAdd(&g_phase_clocks[29], (unsigned long long)clock64() - pc_c0);
    atomicAdd(&g_phase_clocks[30], t1 - pc_t0);
  }
#endif
}

// ---------------------------------------------------------------------------
// K1 + K3 of the CONDENSED fast mode (osc::CoreC, osc_condensed.cuh): Cholesky of M, G =
// M^-1 [B Jc], the (u, z)-only QP with P' = G' Hd G + R, OSQP-style scaling / ADMM on n' = nu +
// 3 nc variables with K^-1 register resident.  One warp per environment, persistent CTAs,
// dynamic work counter, the environment's record landed by TMA bulk copies.
// ---------------------------------------------------------------------------
struct CondArgs {
  const double *M, *C, *J, *mask, *Hdv, *fdv;
  double* state;  // [n_envs][WorkspaceC::STATE]: previous w, dual, rho, flag
  double *torque, *sol_x, *sol_y, *pri_res, *dua_res, *rho;
  int *iters, *status;
  unsigned* counter;
  unsigned base;
  const int* order;
  int n_envs;
};

constexpr int cond_max_regs(int warps) {
  const int r = (65536 / (warps * 32)) / 8 * 8;
  return r > 255 ? 255 : r;
}
template <class D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(cond_max_regs(WARPS))
condensed_kernel(const __grid_constant__ Params p, const CondArgs a) {
  using WS = WorkspaceC<D>;
  using CC = CoreC<D>;
  constexpr int NV = D::NV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WS* wsb = reinterpret_cast<WS*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + WARPS * sizeof(WS));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WS& w = wsb[warp];
  uint64_t* bar = &bars[warp];
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  constexpr uint32_t kBytes = sizeof(typename WS::Stage);
  static_assert(sizeof(typename WS::Stage) ==
                    sizeof(double) * (2 * NV * NV + D::NZ * NV + 2 * NV + D::NC + WS::STATE),
                "the landing stage is exactly the seven bulk copies");
  // (Taking the environments in lockstep batches per CTA, so that the warps share the fetched
  //  instruction lines, was measured: slower, 0.514 -> 0.562 ms -- profiles/r2_experiments.md.)
  uint32_t parity = 0;
  for (;;) {
    int env = 0;
    if (lane == 0) {
      const int t = (int)(atomicAdd(a.counter, 1u) - a.base);
      env = (a.order && t < a.n_envs) ? a.order[t] : t;
      if (env < a.n_envs) {
        fence_proxy_async();  // the stage was written through the generic proxy (L over M)
        mbar_expect_tx(bar, kBytes);
        bulk_g2s(w.m(), a.M + (size_t)env * NV * NV, sizeof(double) * NV * NV, bar);
        bulk_g2s(w.in.H, a.Hdv + (size_t)env * NV * NV, sizeof(w.in.H), bar);
        bulk_g2s(w.jc(), a.J + ((size_t)env * D::S + D::JC0) * NV, sizeof(double) * D::NZ * NV, bar);
        bulk_g2s(w.in.Cv, a.C + (size_t)env * NV, sizeof(w.in.Cv), bar);
        bulk_g2s(w.in.fv, a.fdv + (size_t)env * NV, sizeof(w.in.fv), bar);
        bulk_g2s(w.in.maskv, a.mask + (size_t)env * D::NC, sizeof(w.in.maskv), bar);
        bulk_g2s(w.in.st, a.state + (size_t)env * WS::STATE, sizeof(w.in.st), bar);
      }
    }
    env = __shfl_sync(0xffffffffu, env, 0);
    if (env >= a.n_envs) break;
    __syncwarp();
    mbar_wait(bar, parity);
    parity ^= 1;
    const Result r = CC::step(w, p, lane, a.sol_x + (size_t)env * D::N, a.sol_y + (size_t)env * D::M,
                              a.torque + (size_t)env * D::NU, a.state + (size_t)env * WS::STATE);
    if (lane == 0) {
      a.iters[env] = r.iter;
      a.status[env] = r.status;
      a.pri_res[env] = r.pri_res;
      a.dua_res[env] = r.dua_res;
      a.rho[env] = r.rho;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Longest-processing-time-first order for solve_kernel3's work counter.  The kernel's
// duration is bounded below by its slowest environment; iteration counts of warm-started
// control steps are persistent (the ill-conditioned environments stay so), so every few
// steps the environments are bucketed by their last iteration count (counting sort, one
// CTA) and the solve kernel hands out the expensive ones first instead of meeting one of
// them in its last wave.
// ---------------------------------------------------------------------------
constexpr int kOrderBuckets = 64;
__global__ void __launch_bounds__(1024) order_kernel(const int* __restrict__ iters, int n,
                                                     int* __restrict__ order) {
  __shared__ int cursor[kOrderBuckets];
  const int lane = threadIdx.x & 31;
  // most environments share one bucket: the lanes of a warp that do are counted / placed with
  // ONE shared-memory atomic (match.any), not one each on the same word
  auto claim = [&](int e) -> int {
    const bool valid = e < n;
    int b = -1;
    if (valid) {
      b = iters[e] >> 3;
      b = b < kOrderBuckets ? b : kOrderBuckets - 1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (valid && lane == leader) base = atomicAdd(&cursor[b], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return valid ? base + __popc(peers & ((1u << lane) - 1u)) : -1;
  };
  if (threadIdx.x < kOrderBuckets) cursor[threadIdx.x] = 0;
  __syncthreads();
  const int rounds = (n + blockDim.x - 1) / blockDim.x;  // same trip count for every thread
  for (int r = 0; r < rounds; ++r) claim(r * blockDim.x + threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) {  // first slot of every bucket, largest iteration counts first
    int at = 0;
    for (int b = kOrderBuckets - 1; b >= 0; --b) {
      const int c = cursor[b];
      cursor[b] = at;
      at += c;
    }
  }
  __syncthreads();
  for (int r = 0; r < rounds; ++r) {
    const int e = r * blockDim.x + threadIdx.x;
    const int pos = claim(e);
    if (pos >= 0) order[pos] = e;
  }
}

// ---------------------------------------------------------------------------
// The step before the hot path: task-space PD targets and contact masks (elementwise,
// HBM-bound; one thread per (environment, site) / per environment)
// ---------------------------------------------------------------------------
struct PdGains {
  double kp_lin[kMaxSites], kd_lin[kMaxSites], kp_ang[kMaxSites], kd_ang[kMaxSites];
};

__global__ void __launch_bounds__(256)
targets_pd_kernel(const __grid_constant__ PdGains g, const double* __restrict__ pos,
                  const double* __restrict__ quat, const double* __restrict__ vel,
                  const double* __restrict__ angvel, const double* __restrict__ pos_des,
                  const double* __restrict__ quat_des, const double* __restrict__ vel_des,
                  const double* __restrict__ angvel_des, double* __restrict__ targets, int ns,
                  long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int site = (int)(i % ns);
    const double* p = pos + 3 * i;
    const double* pd = pos_des + 3 * i;
    const double* v = vel + 3 * i;
    const double* w = angvel + 3 * i;
    const double* q = quat + 4 * i;
    const double* qd = quat_des + 4 * i;
    // rotation error = vec(q_des (x) conj(q))   (standing.cc:151)
    const double w1 = qd[0], x1 = qd[1], y1 = qd[2], z1 = qd[3];
    const double w2 = q[0], x2 = -q[1], y2 = -q[2], z2 = -q[3];
    const double ex = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
    const double ey = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
    const double ez = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
    const double e3[3] = {ex, ey, ez};
    double* t = targets + 6 * i;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double vd = vel_des ? vel_des[3 * i + k] : 0.0;
      const double wd = angvel_des ? angvel_des[3 * i + k] : 0.0;
      t[k] = g.kp_lin[site] * (pd[k] - p[k]) + g.kd_lin[site] * (vd - v[k]);
      t[3 + k] = g.kp_ang[site] * e3[k] + g.kd_ang[site] * (wd - w[k]);
    }
  }
}

// examples/walter_sr_true_tumbling_mjjoint.cc:695-802, 873-973, 1001-1019: the tumbling
// driver's target laws, one thread per (environment, site row)
__global__ void __launch_bounds__(256)
targets_walter_tumbling_kernel(const osc_walter_tumbling_gains g, const double* __restrict__ sa,
                               const double* __restrict__ sp, const double* __restrict__ s0,
                               const double* __restrict__ tz, const double* __restrict__ tp,
                               const double* __restrict__ t0, double time, double inv_dt,
                               double* __restrict__ targets, int ns, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e = i / ns;
    const int row = (int)(i - e * ns);
    double t[6] = {0, 0, 0, 0, 0, 0};
    if (row >= 1 && row <= 4) {
      const long long k = 4 * e + (row - 1);
      const double vel = (sa[k] - sp[k]) * inv_dt;
      t[4] = g.shin_kp * ((s0[k] + g.shin_rate * time) - sa[k]) + g.shin_kv * (g.shin_rate - vel);
    } else if (row >= 5 && row <= 8) {
      const long long k = 4 * e + (row - 5);
      const double vel = (tz[k] - tp[k]) * inv_dt;
      t[2] = g.thigh_kp * ((t0[k] - 0.0 + g.thigh_height_offset) - tz[k]) + g.thigh_kv * (g.thigh_rate - vel);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) targets[6 * i + c] = t[c];
  }
}

struct ContactIds {
  int id[kMaxNu];         // nc <= 16 listed geoms
  unsigned bits[kMaxNu];  // mask entries a contact on listed geom j raises
  int nc;
};

__global__ void __launch_bounds__(256)
contact_mask_kernel(const __grid_constant__ ContactIds ids, const int* __restrict__ pairs,
                    const int* __restrict__ ncon, int max_con, double* __restrict__ mask,
                    int n_envs) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_envs; e += gridDim.x * blockDim.x) {
    int n = ncon[e];
    n = n < 0 ? 0 : (n > max_con ? max_con : n);
    unsigned bits = 0;
    const int2* pr = reinterpret_cast<const int2*>(pairs) + (size_t)e * max_con;
    for (int k = 0; k < n; ++k) {
      const int2 gp = pr[k];
      for (int c = 0; c < ids.nc; ++c)
        if (gp.x == ids.id[c] || gp.y == ids.id[c]) bits |= ids.bits[c];
    }
    for (int c = 0; c < ids.nc; ++c) mask[(size_t)e * ids.nc + c] = (bits >> c) & 1u ? 1.0 : 0.0;
  }
}

// ---------------------------------------------------------------------------
// K4 (off the hot path): the "all-gather" of torques and statistics as PEER STORES.  Every
// rank owns a gathered buffer [world][n_envs][nu] (+ [world][kGatherStats] statistics) that
// its peers have mapped (CUDA IPC between processes, plain pointers inside one process); a
// rank pushes its own slice into every mapped copy -- 16-byte stores over NVLink / NVSwitch,
// no staging, no collective call, no NCCL kernel.  Block 0 reduces the rank's statistics
// (count, solved, iteration sum / max, residual maxima) and pushes them the same way.
// ---------------------------------------------------------------------------
constexpr int kGatherStats = 8;
constexpr int kMaxPeers = 16;
struct GatherArgs {
  double* peer[kMaxPeers];   // base of every rank's gathered slab (own included)
  const double* torque;      // local torques [n_envs][nu]
  const int *iters, *status;
  const double *pri_res, *dua_res;
  const int* reinits;
  size_t slice;              // doubles per rank in the torque part = n_envs * nu
  size_t stats_off;          // offset of the statistics part in a slab
  int rank, world, n_envs;
  double step;
};

// FP32 transport of the task Jacobian (opt-in, osc_step_host_j32): rows [r0, r1) of every
// environment of a chunk, float -> double, into the FP64 input buffer the kernels read.
__global__ void __launch_bounds__(256)
widen_rows_kernel(const float* __restrict__ src, double* __restrict__ dst, int n_envs,
                  int row_elems /* s * nv */, int off /* r0 * nv */, int len /* (r1 - r0) * nv */) {
  const size_t total = (size_t)n_envs * len;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i / len, k = i - e * len;
    dst[e * row_elems + off + k] = (double)src[e * row_elems + off + k];
  }
}

constexpr int kGatherThreads = 1024;
__global__ void __launch_bounds__(kGatherThreads) gather_push_kernel(const __grid_constant__ GatherArgs a) {
  if (blockIdx.x == 0) {
    // statistics of this rank
    double cnt = 0, solved = 0, isum = 0, imax = 0, pmax = 0, dmax = 0;
#pragma unroll 8
    for (int e = threadIdx.x; e < a.n_envs; e += blockDim.x) {
      cnt += 1.0;
      solved += a.status[e] == 1 ? 1.0 : 0.0;
      const double it = (double)a.iters[e];
      isum += it;
      imax = it > imax ? it : imax;
      const double pr = a.pri_res[e], du = a.dua_res[e];
      pmax = pr > pmax ? pr : pmax;
      dmax = du > dmax ? du : dmax;
    }
    __shared__ double red[6][kGatherThreads / 32];
    double v[6] = {cnt, solved, isum, imax, pmax, dmax};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v[q], o);
        v[q] = q < 3 ? v[q] + t : (t > v[q] ? t : v[q]);
      }
      if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < kGatherStats) {
      double r = 0.0;
      const int q = threadIdx.x;
      if (q < 6) {
        for (int w = 0; w < kGatherThreads / 32; ++w) r = q < 3 ? r + red[q][w] : (red[q][w] > r ? red[q][w] : r);
      } else if (q == 6) {
        r = (double)*a.reinits;
      } else {
        r = a.step;
      }
      for (int p = 0; p < a.world; ++p)
        a.peer[p][a.stats_off + (size_t)a.rank * kGatherStats + q] = r;
    }
    return;
  }
  const size_t pairs = a.slice / 2;  // nu is even: 16-byte elements
  const double2* src = reinterpret_cast<const double2*>(a.torque);
  for (size_t i = (size_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x; i < pairs;
       i += (size_t)(gridDim.x - 1) * blockDim.x) {
    const double2 v = src[i];
#pragma unroll 4
    for (int p = 0; p < a.world; ++p)
      reinterpret_cast<double2*>(a.peer[p] + (size_t)a.rank * a.slice)[i] = v;
  }
}

// ---------------------------------------------------------------------------
// Self-test of the warp primitives osc_core3.cuh is written against (osc_warp.cuh): the
// device readings of the shuffles, reductions and the DMMA tile product, to be compared
// with the host emulation the CPU tests run the same core on.
// in: [18][32] (rows 0-15: non-negative values for max16 / row 0 also for the shuffles and
// the sum; row 16: A fragment, row 17: B fragment)
// out: max16[16], sum, pad, xchg16[32], group4(r=2)[32], d0[32], d1[32], maxn<8>[8] (rows 0-7)
// ---------------------------------------------------------------------------
__global__ void warp_selftest_kernel(const double* __restrict__ in, double* __restrict__ out) {
  __shared__ double scratch[16];
  const int lane0 = threadIdx.x & 31;
  Var<double> m[16], a, b, d0, d1, x, g;
#pragma unroll
  for (int q = 0; q < 16; ++q) m[q][lane0] = in[32 * q + lane0];
  a[lane0] = in[32 * 16 + lane0];
  b[lane0] = in[32 * 17 + lane0];
  d0[lane0] = 1.0;
  d1[lane0] = -1.0;
  double r16[16], r8[8];
  Var<double> row0 = m[0];
  Var<double> m8[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) m8[q] = m[q];
  Warp::max16(m, r16, scratch, lane0);
  __syncwarp();
  Warp::maxn<8>(m8, r8, scratch, lane0);
  const double sm = Warp::sum(row0);
  Warp::xchg16(x, row0);
  Warp::group4(g, row0, 2);
  Warp::mma884(d0, d1, a, b);
  if (lane0 < 16) out[lane0] = r16[lane0];
  if (lane0 == 0) out[16] = sm;
  out[18 + lane0] = x[lane0];
  out[50 + lane0] = g[lane0];
  out[82 + lane0] = d0[lane0];
  out[114 + lane0] = d1[lane0];
  if (lane0 < 8) out[146 + lane0] = r8[lane0];
}

// ---------------------------------------------------------------------------
// FP64 FMA peak
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4,
         a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace osc

// ===========================================================================
// C-ABI
// ===========================================================================
struct osc_handle {
  osc_robot_spec spec;
  osc_settings settings;
  osc::Params params;
  osc::Shape shape;
  int n_envs, device, sm_count;
  bool setup_done;
  // sizes per environment (doubles)
  int nv, nu, nc, ns, n, m, s, state;
  // device buffers; the six inputs are carved from one slab (dIn) so that a small batch can be
  // uploaded with a single copy (osc_step_host's few-robot path)
  double *dIn;
  size_t in_doubles;                  // slab size
  size_t in_off[6];                   // M, C, J, bias, targets, mask (doubles)
  double *hIn, *hTq;                  // pinned staging of that path (allocated on first use)
  double *dM, *dC, *dJ, *dBias, *dTargets, *dMask;
  double *dH, *dF, *dState, *dScal;
  double* dStateC;  // condensed fast mode: previous w, dual, rho, flag per environment
  double *dTorque, *dX, *dY, *dPri, *dDua, *dRho;
  int *dIters, *dStatus, *dCounter;
  std::vector<unsigned> ctr_base;  // first ticket of the next launch on every work counter
  int* dOrder;      // solve order of the resident full-batch step (order_kernel)
  int order_age;    // osc_step calls since it was rebuilt
  bool use_order;   // set by osc_step around its full-batch solve launch
  // inputs actually read by the kernels (own buffers unless osc_bind_device_inputs)
  const double *iM, *iC, *iJ, *iBias, *iTargets, *iMask;
  // pipelined host path: copy stream + per-chunk events and work counters
  cudaStream_t copy_stream;
  std::vector<cudaEvent_t> chunk_ev;
  cudaEvent_t fence_ev;
  int n_counters;
  // row ranges [first, last) of the task Jacobian the kernels read: rows with a non-zero
  // objective weight plus the contact rows (= contact_jacobian'); osc_step_host uploads these
  std::vector<std::pair<int, int>> j_rows;
  size_t host_h2d_bytes, host_d2h_bytes;  // traffic of the last osc_step_host
  cudaEvent_t timing_mid;  // set while a timed step is being recorded: scale | solve boundary
  int build_grid_max;   // resident CTAs of build_qp_kernel on the device
  bool kernels_ready;   // function attributes of the scale / solve kernels are set
  bool build_ready;     // ... of the build kernel (+ its resident grid size)
  bool cond_ready;      // ... of the condensed kernel
  float* dJ32;          // FP32 landing buffer of the task Jacobian (osc_step_host_j32), lazily allocated
  bool fused_ready;     // ... of the fused build + equilibration kernel
  bool scale_ready;     // ... of the stand-alone equilibration kernel
  bool fuse_build;      // osc_step runs build_scale_kernel3 instead of build + scale (default)
  bool fuse_now;        // set by osc_step around its launch_solve call
  // optional per-kernel timing
  bool timing;
  std::vector<cudaEvent_t> ev;  // 3 events per recorded step
  size_t ev_used;
  long long launches;
  // multi-GPU gather by peer stores (osc_gather_*): own slab + the mapped slabs of the peers
  int g_rank, g_world;
  double* g_slab;                       // [world][n_envs][nu] torques + [world][8] statistics
  double* g_peer[osc::kMaxPeers];
  bool g_ipc_opened[osc::kMaxPeers];
  bool g_attached;
  long long g_steps;
  std::string err;
};

namespace {
thread_local std::string g_create_err;

#define OSC_CUDA(h, call)                                                         \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);              \
      return OSC_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

template <class D>
int launch_build(osc_handle* h, cudaStream_t st, int env0, int n) {
  constexpr int threads = osc::kBuildWarps * 32;
  constexpr int kRows = 4 * ((D::S + 3) / 4);
  const size_t smem = osc::kBuildWarps * osc::kBuildStages * (sizeof(osc::BuildStage<D>) + sizeof(uint64_t)) +
                      kRows * (sizeof(double) + sizeof(int));
  auto kern = osc::build_qp_kernel<D>;
  if (!h->build_ready) {
    OSC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OSC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    h->build_grid_max = h->sm_count * per_sm;
    h->build_ready = true;
  }
  int grid = h->build_grid_max;
  const int need = (n + osc::kBuildWarps - 1) / osc::kBuildWarps;
  if (grid > need) grid = need;
  const size_t e = (size_t)env0;
  kern<<<grid, threads, smem, st>>>(h->params, h->iJ + e * D::S * D::NV, h->iBias + e * D::S,
                                    h->iTargets + e * D::S, h->dH + e * D::NV * D::NV,
                                    h->dF + e * D::NV, n);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

template <class D>
int launch_init_state(osc_handle* h, cudaStream_t st) {
  const int threads = 256;
  int grid = (h->n_envs + 7) / 8;
  if (grid > h->sm_count * 8) grid = h->sm_count * 8;
  osc::init_state_kernel<D><<<grid, threads, 0, st>>>(h->dState, h->dF, h->dH, h->iM, h->iJ,
                                                      h->params.rho0, h->n_envs);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

template <class D>
int launch_reset(osc_handle* h, cudaStream_t st) {
  const int threads = 256;
  const size_t total = (size_t)h->n_envs * (D::N + 2 * D::M);
  int grid = (int)((total + threads - 1) / threads);
  if (grid > h->sm_count * 8) grid = h->sm_count * 8;
  osc::reset_warm_kernel<D><<<grid, threads, 0, st>>>(h->dState, h->n_envs);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

// solve_kernel3: at most 8 warps per CTA (255 registers per thread), fewer when 8 workspaces
// (landing stage included) do not fit in 227 KB of shared memory
// (Registers are allocated to a CTA in units of four warps: 9 or 10 warps at 224 / 200
// registers do not launch.  The next step after 8 warps at 255 registers is 12 at 168.)
#ifndef OSC_SOLVE_WARPS
#define OSC_SOLVE_WARPS 8
#endif
template <class D>
constexpr int solve3_warps() {
  constexpr int fit = (int)((227 * 1024 - 128) / sizeof(osc::Workspace3<D>));
  return fit > OSC_SOLVE_WARPS ? OSC_SOLVE_WARPS : fit;
}
// scale_kernel3: 12 warps per CTA at 168 registers when two lanes share a row (16 warps at 128
// registers spill a little and measured the same 0.17 ms); rows held by one lane need more
// registers per thread: 8 warps
template <class D>
constexpr int scale3_warps() { return (2 * D::NV <= 32) ? 12 : 8; }

template <class D>
int launch_scale3(osc_handle* h, cudaStream_t st, int env0, int n, int counter) {
  constexpr int WARPS = scale3_warps<D>();
  const size_t smem = WARPS * sizeof(osc::RuizWorkspace<D>) + WARPS * sizeof(uint64_t);
  auto kern = osc::scale_kernel3<D, WARPS>;
  if (!h->scale_ready) {
    OSC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->scale_ready = true;
  }
  int grid = h->sm_count;
  const int need = (n + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  const int slot = h->n_counters + 1 + counter;  // second bank of work counters
  const size_t e = (size_t)env0;
  osc::ScaleArgs a;
  a.M = h->iM + e * D::NV * D::NV; a.J = h->iJ + e * D::S * D::NV;
  a.Hdv = h->dH + e * D::NV * D::NV; a.fdv = h->dF + e * D::NV;
  a.bias = a.targets = nullptr; a.Hdv_out = a.fdv_out = nullptr;
  a.state = h->dState + e * D::STATE;
  a.scal = h->dScal + e * osc::Core3<D>::SCAL;
  a.counter = reinterpret_cast<unsigned*>(h->dCounter) + slot;
  a.base = h->ctr_base[slot];
  a.n_envs = n;
  kern<<<grid, WARPS * 32, smem, st>>>(h->params, a);
  OSC_CUDA(h, cudaGetLastError());
  h->ctr_base[slot] += (unsigned)n + 2u * (unsigned)(grid * WARPS);
  h->launches++;
  return OSC_OK;
}

// fused objective build + equilibration (build_scale_kernel3): same warps per CTA as the
// equilibration alone; the landing stage grows by the task Jacobian
template <class D>
int launch_build_scale3(osc_handle* h, cudaStream_t st, int env0, int n, int counter) {
  constexpr int WARPS = scale3_warps<D>();
  constexpr int KSTEPS = (D::S + 3) / 4;
  constexpr size_t smem = WARPS * sizeof(osc::BuildRuizWorkspace<D>) + WARPS * sizeof(uint64_t) +
                          4 * KSTEPS * (sizeof(double) + sizeof(int));
  static_assert(smem <= 227 * 1024, "shared memory per CTA");
  auto kern = osc::build_scale_kernel3<D, WARPS>;
  if (!h->fused_ready) {
    OSC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->fused_ready = true;
  }
  int grid = h->sm_count;
  const int need = (n + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  const int slot = h->n_counters + 1 + counter;  // second bank of work counters
  const size_t e = (size_t)env0;
  osc::ScaleArgs a;
  a.M = h->iM + e * D::NV * D::NV; a.J = h->iJ + e * D::S * D::NV;
  a.Hdv = nullptr; a.fdv = nullptr;
  a.bias = h->iBias + e * D::S; a.targets = h->iTargets + e * D::S;
  a.Hdv_out = h->dH + e * D::NV * D::NV; a.fdv_out = h->dF + e * D::NV;
  a.state = h->dState + e * D::STATE;
  a.scal = h->dScal + e * osc::Core3<D>::SCAL;
  a.counter = reinterpret_cast<unsigned*>(h->dCounter) + slot;
  a.base = h->ctr_base[slot];
  a.n_envs = n;
  kern<<<grid, WARPS * 32, smem, st>>>(h->params, a);
  OSC_CUDA(h, cudaGetLastError());
  h->ctr_base[slot] += (unsigned)n + 2u * (unsigned)(grid * WARPS);
  h->launches++;
  return OSC_OK;
}

template <class D, int WARPS>
int launch_solve3w(osc_handle* h, cudaStream_t st, int env0, int n, int counter) {
  const size_t smem = WARPS * sizeof(osc::Workspace3<D>) + WARPS * sizeof(uint64_t);
  auto kern = osc::solve_kernel3<D, WARPS>;
  if (!h->kernels_ready)
    OSC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = h->sm_count;
  const int need = (n + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  const size_t e = (size_t)env0;
  osc::SolveArgs a;
  a.M = h->iM + e * D::NV * D::NV; a.C = h->iC + e * D::NV; a.J = h->iJ + e * D::S * D::NV;
  a.mask = h->iMask + e * D::NC; a.Hdv = h->dH + e * D::NV * D::NV; a.fdv = h->dF + e * D::NV;
  a.state = h->dState + e * D::STATE;
  a.torque = h->dTorque + e * D::NU; a.sol_x = h->dX + e * D::N; a.sol_y = h->dY + e * D::M;
  a.pri_res = h->dPri + e; a.dua_res = h->dDua + e; a.rho = h->dRho + e;
  a.iters = h->dIters + e; a.status = h->dStatus + e;
  a.counter = reinterpret_cast<unsigned*>(h->dCounter) + counter;
  a.base = h->ctr_base[counter];
  a.order = (h->use_order && env0 == 0 && n == h->n_envs) ? h->dOrder : nullptr;
  a.reinits = h->dCounter + h->n_counters;
  a.scal = h->dScal + e * osc::Core3<D>::SCAL;
  a.n_envs = n;
  kern<<<grid, WARPS * 32, smem, st>>>(h->params, a);
  OSC_CUDA(h, cudaGetLastError());
  h->ctr_base[counter] += (unsigned)n + 2u * (unsigned)(grid * WARPS);
  h->launches++;
  return OSC_OK;
}

template <class D>
int launch_solve3(osc_handle* h, cudaStream_t st, int env0, int n, int counter) {
  int rc = h->fuse_now ? launch_build_scale3<D>(h, st, env0, n, counter)
                       : launch_scale3<D>(h, st, env0, n, counter);
  if (rc) return rc;
  if (h->timing_mid) OSC_CUDA(h, cudaEventRecord(h->timing_mid, st));
  return launch_solve3w<D, solve3_warps<D>()>(h, st, env0, n, counter);
}

template <class D>
int launch_solve(osc_handle* h, cudaStream_t st, int env0, int n, int counter) {
  return launch_solve3<D>(h, st, env0, n, counter);
}

// condensed_kernel: warps per CTA.  12 at 168 registers is the default; the developer knob
// OSC_B200_COND_WARPS (environment, read once) selects 8 / 10 / 12 / 15 for occupancy experiments.
template <class D, int WARPS>
int launch_condensed_w(osc_handle* h, cudaStream_t st, int n) {
  static_assert(WARPS * sizeof(osc::WorkspaceC<D>) + 256 <= 227 * 1024, "shared memory per CTA");
  const size_t smem = WARPS * sizeof(osc::WorkspaceC<D>) + WARPS * sizeof(uint64_t);
  auto kern = osc::condensed_kernel<D, WARPS>;
  if (!h->cond_ready) {
    OSC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->cond_ready = true;
  }
  int grid = h->sm_count;
  const int need = (n + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  const int slot = 2 * h->n_counters + 1;  // last work counter: the condensed launches
  osc::CondArgs a;
  a.M = h->iM; a.C = h->iC; a.J = h->iJ; a.mask = h->iMask; a.Hdv = h->dH; a.fdv = h->dF;
  a.state = h->dStateC;
  a.torque = h->dTorque; a.sol_x = h->dX; a.sol_y = h->dY;
  a.pri_res = h->dPri; a.dua_res = h->dDua; a.rho = h->dRho;
  a.iters = h->dIters; a.status = h->dStatus;
  a.counter = reinterpret_cast<unsigned*>(h->dCounter) + slot;
  a.base = h->ctr_base[slot];
  a.order = h->use_order ? h->dOrder : nullptr;
  a.n_envs = n;
  kern<<<grid, WARPS * 32, smem, st>>>(h->params, a);
  OSC_CUDA(h, cudaGetLastError());
  h->ctr_base[slot] += (unsigned)n + (unsigned)(grid * WARPS);
  h->launches++;
  return OSC_OK;
}

template <class D>
int launch_condensed(osc_handle* h, cudaStream_t st, int n) {
  static const int warps = [] {
    const char* e = std::getenv("OSC_B200_COND_WARPS");
    return e ? std::atoi(e) : 12;
  }();
  switch (warps) {
    case 8: return launch_condensed_w<D, 8>(h, st, n);
    case 10: return launch_condensed_w<D, 10>(h, st, n);
    case 15: return launch_condensed_w<D, 15>(h, st, n);
    default: return launch_condensed_w<D, 12>(h, st, n);
  }
}

// ---- device-side kinematics / dynamics (the step before the hot path) ------------------------
template <class D>
int launch_kinematics(osc_handle* h, const osc_kin_model* m, const double* qpos,
                             const double* qvel, cudaStream_t st) {
  const size_t smem = osc::kKinWarps * sizeof(osc::KinWorkspace<D>);
  auto kern = osc::kinematics_kernel<D>;
  OSC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (h->n_envs + osc::kKinWarps - 1) / osc::kKinWarps;
  if (grid > h->sm_count * 2) grid = h->sm_count * 2;
  kern<<<grid, osc::kKinWarps * 32, smem, st>>>(*m, qpos, qvel, h->dM, h->dC, h->dJ, h->dBias,
                                               h->n_envs);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}


#define OSC_DISPATCH(h, fn, ...)                                                    \
  ((h)->shape == osc::Shape::kWalter ? fn<osc::WalterDims>(__VA_ARGS__)             \
                                     : fn<osc::Go2Dims>(__VA_ARGS__))

int check_handle(const osc_handle* h) { return h ? OSC_OK : OSC_ERR_INVALID; }
}  // namespace

extern "C" {

int osc_default_settings(osc_settings* s) {
  if (!s) return OSC_ERR_INVALID;
  osc::default_settings(s);
  return OSC_OK;
}

const char* osc_last_error(const osc_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int osc_create(const osc_robot_spec* spec, const osc_settings* settings, int n_envs, int device,
               osc_handle** out) {
  if (!spec || !out || n_envs <= 0) {
    g_create_err = "osc_create: null argument or n_envs <= 0";
    return OSC_ERR_INVALID;
  }
  *out = nullptr;
  const osc::Shape shape = osc::shape_of(*spec);
  if (shape == osc::Shape::kNone) {
    g_create_err = "osc_create: unsupported robot shape (compiled: nv,nu,nc,ns = 14,8,8,17 | 18,12,4,5)";
    return OSC_ERR_INVALID;
  }
  osc_settings st;
  if (settings) st = *settings; else osc::default_settings(&st);
  if (st.max_iter < 1 || st.check_termination < 0 || st.scaling < 0 || st.rho <= 0 ||
      st.sigma <= 0 || st.alpha <= 0 || st.alpha >= 2 || st.eps_prim_inf <= 0 ||
      st.eps_dual_inf <= 0 || st.eps_abs < 0 || st.eps_rel < 0 ||
      (st.adaptive_rho && st.adaptive_rho_tolerance < 1.0) || st.adaptive_rho_interval < 0) {
    // the ranges of OSQP 0.6.3 validate_settings (auxil.c) for the fields osc_settings carries;
    // eps_abs == eps_rel == 0 stays legal here: it is the fixed-iteration-budget mode
    g_create_err = "osc_create: invalid settings (OSQP validate_settings ranges)";
    return OSC_ERR_INVALID;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count) {
    g_create_err = std::string("osc_create: no usable CUDA device (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                   "); this library has no CPU fallback";
    return OSC_ERR_CUDA;
  }
  osc_handle* h = new (std::nothrow) osc_handle();
  if (!h) return OSC_ERR_ALLOC;
  h->spec = *spec;
  h->settings = st;
  h->params = osc::make_params(*spec, st);
  h->shape = shape;
  h->n_envs = n_envs;
  h->device = device;
  h->setup_done = false;
  h->launches = 0;
  h->nv = spec->nv; h->nu = spec->nu; h->nc = spec->nc; h->ns = spec->ns;
  h->n = spec->nv + spec->nu + 3 * spec->nc;
  h->m = spec->nv + 4 * spec->nc + h->n;
  h->s = 6 * spec->ns;
  h->state = shape == osc::Shape::kWalter ? osc::WalterDims::STATE : osc::Go2Dims::STATE;
  auto fail = [&](cudaError_t ce, const char* what) {
    g_create_err = std::string(what) + ": " + cudaGetErrorString(ce);
    osc_destroy(h);
    return ce == cudaErrorMemoryAllocation ? OSC_ERR_ALLOC : OSC_ERR_CUDA;
  };
  cudaError_t ce;
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail(ce, "cudaSetDevice");
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(ce, "cudaGetDeviceProperties");
  if (prop.major != 10 || prop.minor != 0) {
    // architecture-specific SASS (sm_100a) has no forward compatibility: CC 10.3 / 12.x devices
    // would fail later with "no kernel image"
    g_create_err = "osc_create: kernels are built for sm_100a (compute capability 10.0) only; device is " +
                   std::to_string(prop.major) + "." + std::to_string(prop.minor);
    osc_destroy(h);
    return OSC_ERR_CUDA;
  }
  h->sm_count = prop.multiProcessorCount;
  const size_t N = (size_t)n_envs;
  {
    const size_t sz[6] = {N * h->nv * h->nv, N * h->nv, N * h->s * h->nv, N * h->s, N * h->s,
                          N * h->nc};
    size_t off = 0;
    for (int k = 0; k < 6; ++k) {
      h->in_off[k] = off;
      off += (sz[k] + 31) & ~(size_t)31;  // 256-byte aligned regions
    }
    h->in_doubles = off;
    if ((ce = cudaMalloc((void**)&h->dIn, off * sizeof(double))) != cudaSuccess) return fail(ce, "cudaMalloc");
    if ((ce = cudaMemset(h->dIn, 0, off * sizeof(double))) != cudaSuccess) return fail(ce, "cudaMemset");
    h->dM = h->dIn + h->in_off[0]; h->dC = h->dIn + h->in_off[1]; h->dJ = h->dIn + h->in_off[2];
    h->dBias = h->dIn + h->in_off[3]; h->dTargets = h->dIn + h->in_off[4];
    h->dMask = h->dIn + h->in_off[5];
  }
  struct { double** p; size_t n; } bufs[] = {
      {&h->dH, N * h->nv * h->nv}, {&h->dF, N * h->nv}, {&h->dState, N * h->state},
      {&h->dTorque, N * h->nu}, {&h->dX, N * h->n}, {&h->dY, N * h->m},
      {&h->dPri, N}, {&h->dDua, N}, {&h->dRho, N}, {&h->dScal, N * (size_t)(h->n + h->m + 2)}};
  for (auto& b : bufs) {
    if ((ce = cudaMalloc((void**)b.p, b.n * sizeof(double))) != cudaSuccess) return fail(ce, "cudaMalloc");
    if ((ce = cudaMemset(*b.p, 0, b.n * sizeof(double))) != cudaSuccess) return fail(ce, "cudaMemset");
  }
  if ((ce = cudaMalloc((void**)&h->dIters, N * sizeof(int))) != cudaSuccess) return fail(ce, "cudaMalloc");
  if ((ce = cudaMalloc((void**)&h->dStatus, N * sizeof(int))) != cudaSuccess) return fail(ce, "cudaMalloc");
  h->n_counters = 64;
  // work counters of the solve launches [0, n), the re-Init count [n], work counters of the
  // scale launches [n + 1, 2n + 1)
  if ((ce = cudaMalloc((void**)&h->dCounter, (2 * h->n_counters + 2) * sizeof(int))) != cudaSuccess) return fail(ce, "cudaMalloc");
  if ((ce = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(ce, "cudaStreamCreate");
  if ((ce = cudaEventCreateWithFlags(&h->fence_ev, cudaEventDisableTiming)) != cudaSuccess) return fail(ce, "cudaEventCreate");
  cudaMemset(h->dCounter, 0, (2 * h->n_counters + 2) * sizeof(int));
  h->ctr_base.assign(2 * h->n_counters + 2, 0u);
  if ((ce = cudaMalloc((void**)&h->dOrder, N * sizeof(int))) != cudaSuccess) return fail(ce, "cudaMalloc");
  h->order_age = -1;  // no order yet
  h->use_order = false;
  cudaMemset(h->dIters, 0, N * sizeof(int));
  cudaMemset(h->dStatus, 0, N * sizeof(int));
  h->iM = h->dM; h->iC = h->dC; h->iJ = h->dJ; h->iBias = h->dBias; h->iTargets = h->dTargets;
  h->iMask = h->dMask;
  {
    const int S = h->s, jc0 = 3 * h->ns - 3 * h->nc, jc1 = 3 * h->ns;
    int first = -1;
    for (int k = 0; k <= S; ++k) {
      const bool live = k < S && (h->params.w_row[k] != 0.0 || (k >= jc0 && k < jc1));
      if (live && first < 0) first = k;
      if (!live && first >= 0) {
        h->j_rows.emplace_back(first, k);
        first = -1;
      }
    }
  }
  h->host_h2d_bytes = h->host_d2h_bytes = 0;
  h->dStateC = nullptr;
  h->g_rank = 0; h->g_world = 0; h->g_slab = nullptr; h->g_attached = false; h->g_steps = 0;
  for (int p = 0; p < osc::kMaxPeers; ++p) { h->g_peer[p] = nullptr; h->g_ipc_opened[p] = false; }
  h->timing = false;
  h->timing_mid = nullptr;
  h->ev_used = 0;
  h->kernels_ready = false;
  h->build_ready = false;
  h->cond_ready = false;
  h->dJ32 = nullptr;
  h->fused_ready = false;
  h->scale_ready = false;
  h->fuse_build = true;
  h->fuse_now = false;
  h->build_grid_max = h->sm_count;
  *out = h;
  return OSC_OK;
}

int osc_destroy(osc_handle* h) {
  if (!h) return OSC_ERR_INVALID;
  cudaSetDevice(h->device);
  double* d[] = {h->dIn, h->dH, h->dF, h->dState,
                 h->dTorque, h->dX, h->dY, h->dPri, h->dDua, h->dRho, h->dScal};
  for (double* p : d) if (p) cudaFree(p);
  for (int p = 0; p < osc::kMaxPeers; ++p)
    if (h->g_ipc_opened[p] && h->g_peer[p]) cudaIpcCloseMemHandle(h->g_peer[p]);
  if (h->g_slab) cudaFree(h->g_slab);
  if (h->dStateC) cudaFree(h->dStateC);
  if (h->dJ32) cudaFree(h->dJ32);
  if (h->hIn) cudaFreeHost(h->hIn);
  if (h->hTq) cudaFreeHost(h->hTq);
  if (h->dIters) cudaFree(h->dIters);
  if (h->dStatus) cudaFree(h->dStatus);
  if (h->dCounter) cudaFree(h->dCounter);
  if (h->dOrder) cudaFree(h->dOrder);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->chunk_ev) cudaEventDestroy(e);
  if (h->fence_ev) cudaEventDestroy(h->fence_ev);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return OSC_OK;
}

int osc_num_envs(const osc_handle* h) { return h ? h->n_envs : OSC_ERR_INVALID; }

int osc_get_device_buffers(osc_handle* h, osc_device_buffers* out) {
  if (check_handle(h) || !out) return OSC_ERR_INVALID;
  out->M = h->dM; out->C = h->dC; out->J = h->dJ; out->bias = h->dBias;
  out->targets = h->dTargets; out->mask = h->dMask;
  out->torque = h->dTorque; out->solution = h->dX; out->dual = h->dY;
  out->iters = h->dIters; out->status = h->dStatus;
  out->pri_res = h->dPri; out->dua_res = h->dDua; out->rho = h->dRho;
  return OSC_OK;
}

int osc_upload(osc_handle* h, const double* M, const double* C, const double* J,
               const double* bias, const double* targets, const double* mask, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  const size_t N = (size_t)h->n_envs, B = sizeof(double);
  if (M) OSC_CUDA(h, cudaMemcpyAsync(h->dM, M, N * h->nv * h->nv * B, cudaMemcpyHostToDevice, st));
  if (C) OSC_CUDA(h, cudaMemcpyAsync(h->dC, C, N * h->nv * B, cudaMemcpyHostToDevice, st));
  if (J) OSC_CUDA(h, cudaMemcpyAsync(h->dJ, J, N * h->s * h->nv * B, cudaMemcpyHostToDevice, st));
  if (bias) OSC_CUDA(h, cudaMemcpyAsync(h->dBias, bias, N * h->s * B, cudaMemcpyHostToDevice, st));
  if (targets) OSC_CUDA(h, cudaMemcpyAsync(h->dTargets, targets, N * h->s * B, cudaMemcpyHostToDevice, st));
  if (mask) OSC_CUDA(h, cudaMemcpyAsync(h->dMask, mask, N * h->nc * B, cudaMemcpyHostToDevice, st));
  return OSC_OK;
}

int osc_setup(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  int rc = OSC_DISPATCH(h, launch_build, h, st, 0, h->n_envs);
  if (rc) return rc;
  rc = OSC_DISPATCH(h, launch_init_state, h, st);
  if (rc) return rc;
  h->setup_done = true;
  return OSC_OK;
}

constexpr int kOrderEvery = 8;

int osc_step(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!h->setup_done) {
    h->err = "osc_step: osc_setup has not been called (set_up_optimization precedes control_loop)";
    return OSC_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  // solve order: rebuilt from the last iteration counts every kOrderEvery steps (batches of
  // more than one wave only; a smaller batch has every environment in flight at once)
  const bool ordered = h->n_envs > h->sm_count * 8;
  if (ordered && (h->order_age < 0 || h->order_age >= kOrderEvery)) {
    osc::order_kernel<<<1, 1024, 0, st>>>(h->dIters, h->n_envs, h->dOrder);
    OSC_CUDA(h, cudaGetLastError());
    h->launches++;
    h->order_age = 0;
  }
  cudaEvent_t* ev = nullptr;
  if (h->timing) {
    if (h->ev_used + 4 > h->ev.size()) {
      for (int i = 0; i < 4; ++i) {
        cudaEvent_t e;
        OSC_CUDA(h, cudaEventCreate(&e));
        h->ev.push_back(e);
      }
    }
    ev = &h->ev[h->ev_used];
    h->ev_used += 4;
    OSC_CUDA(h, cudaEventRecord(ev[0], st));
  }
  int rc = OSC_OK;
  // the objective build runs inside the equilibration kernel (build_scale_kernel3) unless the
  // caller asked for the three-kernel form (osc_set_fused_build: measurements, A/B tests)
  if (!h->fuse_build) rc = OSC_DISPATCH(h, launch_build, h, st, 0, h->n_envs);
  if (rc) return rc;
  if (ev) OSC_CUDA(h, cudaEventRecord(ev[1], st));
  // ev[2] = scale | solve boundary: recorded by launch_solve3 between its two kernels
  h->timing_mid = ev ? ev[2] : nullptr;
  h->use_order = ordered;
  h->fuse_now = h->fuse_build;
  rc = OSC_DISPATCH(h, launch_solve, h, st, 0, h->n_envs, 0);
  h->fuse_now = false;
  h->use_order = false;
  h->timing_mid = nullptr;
  if (rc) return rc;
  if (ordered) h->order_age++;
  h->kernels_ready = true;  // function attributes / occupancy are set from here on
  if (ev) OSC_CUDA(h, cudaEventRecord(ev[3], st));
  return OSC_OK;
}

int osc_set_fused_build(osc_handle* h, int on) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  h->fuse_build = on != 0;
  return OSC_OK;
}

// ---- condensed fast mode ---------------------------------------------------------------------
static int condensed_state(osc_handle* h) {
  if (h->dStateC) return OSC_OK;
  const size_t ss = h->shape == osc::Shape::kWalter ? osc::WorkspaceC<osc::WalterDims>::STATE
                                                    : osc::WorkspaceC<osc::Go2Dims>::STATE;
  const size_t bytes = (size_t)h->n_envs * ss * sizeof(double);
  OSC_CUDA(h, cudaMalloc((void**)&h->dStateC, bytes));
  OSC_CUDA(h, cudaMemset(h->dStateC, 0, bytes));
  return OSC_OK;
}

int osc_step_condensed(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  int rc = condensed_state(h);
  if (rc) return rc;
  const bool ordered = h->n_envs > h->sm_count * 12;
  if (ordered && (h->order_age < 0 || h->order_age >= kOrderEvery)) {
    osc::order_kernel<<<1, 1024, 0, st>>>(h->dIters, h->n_envs, h->dOrder);
    OSC_CUDA(h, cudaGetLastError());
    h->launches++;
    h->order_age = 0;
  }
  cudaEvent_t* ev = nullptr;
  if (h->timing) {
    if (h->ev_used + 4 > h->ev.size()) {
      for (int i = 0; i < 4; ++i) {
        cudaEvent_t e;
        OSC_CUDA(h, cudaEventCreate(&e));
        h->ev.push_back(e);
      }
    }
    ev = &h->ev[h->ev_used];
    h->ev_used += 4;
    OSC_CUDA(h, cudaEventRecord(ev[0], st));
  }
  rc = OSC_DISPATCH(h, launch_build, h, st, 0, h->n_envs);
  if (rc) return rc;
  if (ev) {
    OSC_CUDA(h, cudaEventRecord(ev[1], st));
    OSC_CUDA(h, cudaEventRecord(ev[2], st));  // no separate scale kernel in this mode
  }
  h->use_order = ordered;
  rc = OSC_DISPATCH(h, launch_condensed, h, st, h->n_envs);
  h->use_order = false;
  if (rc) return rc;
  if (ordered) h->order_age++;
  if (ev) OSC_CUDA(h, cudaEventRecord(ev[3], st));
  return OSC_OK;
}

int osc_reset_condensed(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  OSC_CUDA(h, cudaSetDevice(h->device));
  int rc = condensed_state(h);
  if (rc) return rc;
  const size_t ss = h->shape == osc::Shape::kWalter ? osc::WorkspaceC<osc::WalterDims>::STATE
                                                    : osc::WorkspaceC<osc::Go2Dims>::STATE;
  OSC_CUDA(h, cudaMemsetAsync(h->dStateC, 0, (size_t)h->n_envs * ss * sizeof(double),
                              (cudaStream_t)stream));
  h->order_age = -1;
  return OSC_OK;
}

int osc_bind_device_inputs(osc_handle* h, const double* M, const double* C, const double* J,
                           const double* bias, const double* targets, const double* mask) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  h->iM = M ? M : h->dM;
  h->iC = C ? C : h->dC;
  h->iJ = J ? J : h->dJ;
  h->iBias = bias ? bias : h->dBias;
  h->iTargets = targets ? targets : h->dTargets;
  h->iMask = mask ? mask : h->dMask;
  const uintptr_t all = (uintptr_t)h->iM | (uintptr_t)h->iC | (uintptr_t)h->iJ |
                        (uintptr_t)h->iBias | (uintptr_t)h->iTargets | (uintptr_t)h->iMask;
  if (all & 15) {
    h->err = "osc_bind_device_inputs: device pointers must be 16-byte aligned (TMA bulk copies)";
    h->iM = h->dM; h->iC = h->dC; h->iJ = h->dJ; h->iBias = h->dBias; h->iTargets = h->dTargets;
    h->iMask = h->dMask;
    return OSC_ERR_INVALID;
  }
  return OSC_OK;
}

int osc_host_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return OSC_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
    *out = nullptr;
    return OSC_ERR_CUDA;
  }
  return OSC_OK;
}
int osc_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? OSC_OK : OSC_ERR_CUDA; }

int osc_timing_enable(osc_handle* h, int on) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  h->timing = on != 0;
  h->ev_used = 0;
  return OSC_OK;
}

int osc_timing_read(osc_handle* h, osc_kernel_times* out) {
  if (check_handle(h) || !out) return OSC_ERR_INVALID;
  OSC_CUDA(h, cudaSetDevice(h->device));
  const int steps = (int)(h->ev_used / 4);
  double b = 0, sc = 0, s = 0;
  for (int i = 0; i < steps; ++i) {
    OSC_CUDA(h, cudaEventSynchronize(h->ev[4 * i + 3]));
    float t0 = 0, t1 = 0, t2 = 0;
    OSC_CUDA(h, cudaEventElapsedTime(&t0, h->ev[4 * i], h->ev[4 * i + 1]));
    OSC_CUDA(h, cudaEventElapsedTime(&t1, h->ev[4 * i + 1], h->ev[4 * i + 2]));
    OSC_CUDA(h, cudaEventElapsedTime(&t2, h->ev[4 * i + 2], h->ev[4 * i + 3]));
    b += t0;
    sc += t1;
    s += t2;
  }
  out->steps = steps;
  out->build_ms = steps ? (float)(b / steps) : 0.f;
  out->scale_ms = steps ? (float)(sc / steps) : 0.f;
  out->solve_ms = steps ? (float)(s / steps) : 0.f;
  h->ev_used = 0;
  return OSC_OK;
}

int osc_reset_warm_start(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!h->setup_done) {
    h->err = "osc_reset_warm_start: osc_setup has not been called";
    return OSC_ERR_STATE;
  }
  OSC_CUDA(h, cudaSetDevice(h->device));
  return OSC_DISPATCH(h, launch_reset, h, (cudaStream_t)stream);
}

int osc_download(osc_handle* h, double* torque, double* solution, double* dual, int* iters,
                 int* status, double* pri_res, double* dua_res, double* rho, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  const size_t N = (size_t)h->n_envs, B = sizeof(double);
  if (torque) OSC_CUDA(h, cudaMemcpyAsync(torque, h->dTorque, N * h->nu * B, cudaMemcpyDeviceToHost, st));
  if (solution) OSC_CUDA(h, cudaMemcpyAsync(solution, h->dX, N * h->n * B, cudaMemcpyDeviceToHost, st));
  if (dual) OSC_CUDA(h, cudaMemcpyAsync(dual, h->dY, N * h->m * B, cudaMemcpyDeviceToHost, st));
  if (iters) OSC_CUDA(h, cudaMemcpyAsync(iters, h->dIters, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (status) OSC_CUDA(h, cudaMemcpyAsync(status, h->dStatus, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (pri_res) OSC_CUDA(h, cudaMemcpyAsync(pri_res, h->dPri, N * B, cudaMemcpyDeviceToHost, st));
  if (dua_res) OSC_CUDA(h, cudaMemcpyAsync(dua_res, h->dDua, N * B, cudaMemcpyDeviceToHost, st));
  if (rho) OSC_CUDA(h, cudaMemcpyAsync(rho, h->dRho, N * B, cudaMemcpyDeviceToHost, st));
  return OSC_OK;
}

int osc_download_objective(osc_handle* h, double* H_dv, double* f_dv, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  const size_t N = (size_t)h->n_envs, B = sizeof(double);
  if (H_dv) OSC_CUDA(h, cudaMemcpyAsync(H_dv, h->dH, N * h->nv * h->nv * B, cudaMemcpyDeviceToHost, st));
  if (f_dv) OSC_CUDA(h, cudaMemcpyAsync(f_dv, h->dF, N * h->nv * B, cudaMemcpyDeviceToHost, st));
  return OSC_OK;
}

int osc_sync(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  OSC_CUDA(h, cudaSetDevice(h->device));
  OSC_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return OSC_OK;
}

// A few robots (the reference's own use: one robot in a 1 kHz loop): the whole input slab is
// staged in pinned memory and uploaded with ONE copy on the caller's stream, the torques come
// back through pinned memory -- a dozen driver calls fewer than the chunked pipeline, which
// is what such a step's latency consists of.
constexpr size_t kFewRobotsBytes = 128 * 1024;
static int step_host_few(osc_handle* h, const double* M, const double* C, const double* J,
                         const float* J32, const double* bias, const double* targets,
                         const double* mask, double* torque, cudaStream_t st) {
  const size_t B = sizeof(double), N = (size_t)h->n_envs;
  const size_t nv = h->nv, s = h->s, nc = h->nc, nu = h->nu;
  if (!h->hIn) {
    OSC_CUDA(h, cudaHostAlloc((void**)&h->hIn, h->in_doubles * B, cudaHostAllocDefault));
    std::memset(h->hIn, 0, h->in_doubles * B);
  }
  if (!h->hTq) OSC_CUDA(h, cudaHostAlloc((void**)&h->hTq, N * nu * B, cudaHostAllocDefault));
  std::memcpy(h->hIn + h->in_off[0], M, N * nv * nv * B);
  std::memcpy(h->hIn + h->in_off[1], C, N * nv * B);
  for (size_t e = 0; e < N; ++e)  // only the rows the kernels read leave the caller's buffer
    for (const auto& r : h->j_rows) {
      const size_t off = (e * s + (size_t)r.first) * nv, len = (size_t)(r.second - r.first) * nv;
      if (J32) {  // (a few robots: widened while staging)
        for (size_t k = 0; k < len; ++k) h->hIn[h->in_off[2] + off + k] = (double)J32[off + k];
      } else {
        std::memcpy(h->hIn + h->in_off[2] + off, J + off, len * B);
      }
    }
  std::memcpy(h->hIn + h->in_off[3], bias, N * s * B);
  std::memcpy(h->hIn + h->in_off[4], targets, N * s * B);
  std::memcpy(h->hIn + h->in_off[5], mask, N * nc * B);
  OSC_CUDA(h, cudaMemcpyAsync(h->dIn, h->hIn, h->in_doubles * B, cudaMemcpyHostToDevice, st));
  int rc = h->fuse_build ? OSC_OK : OSC_DISPATCH(h, launch_build, h, st, 0, (int)N);
  if (rc) return rc;
  h->fuse_now = h->fuse_build;
  rc = OSC_DISPATCH(h, launch_solve, h, st, 0, (int)N, 0);
  h->fuse_now = false;
  if (rc) return rc;
  OSC_CUDA(h, cudaMemcpyAsync(h->hTq, h->dTorque, N * nu * B, cudaMemcpyDeviceToHost, st));
  h->kernels_ready = true;
  h->host_h2d_bytes = h->in_doubles * B;
  h->host_d2h_bytes = N * nu * B;
  OSC_CUDA(h, cudaStreamSynchronize(st));
  std::memcpy(torque, h->hTq, N * nu * B);
  return OSC_OK;
}

static int step_host_impl(osc_handle* h, const double* M, const double* C, const double* J,
                          const float* J32, const double* bias, const double* targets,
                          const double* mask, double* torque, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!h->setup_done) {
    h->err = "osc_step_host: osc_setup has not been called";
    return OSC_ERR_STATE;
  }
  if (!M || !C || (!J && !J32) || !bias || !targets || !mask || !torque) {
    h->err = "osc_step_host: null host buffer";
    return OSC_ERR_INVALID;
  }
  if (h->iM != h->dM || h->iC != h->dC || h->iJ != h->dJ || h->iBias != h->dBias ||
      h->iTargets != h->dTargets || h->iMask != h->dMask) {
    h->err = "osc_step_host: inputs are bound to caller-owned device memory (osc_bind_device_inputs)";
    return OSC_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  if (h->in_doubles * sizeof(double) <= kFewRobotsBytes)
    return step_host_few(h, M, C, J, J32, bias, targets, mask, torque, st);
  if (J32 && !h->dJ32) {  // FP32 landing buffer; rows that are never uploaded stay zero
    const size_t bytes = (size_t)h->n_envs * h->s * h->nv * sizeof(float);
    OSC_CUDA(h, cudaMalloc((void**)&h->dJ32, bytes));
    OSC_CUDA(h, cudaMemset(h->dJ32, 0, bytes));
  }
  // Software pipeline over chunks of environments: the H2D copy of chunk c+1 (copy stream)
  // overlaps build+solve of chunk c (caller's stream); torques return per chunk.
  const int N = h->n_envs;
  int chunk = 2048;
  int nchunks = (N + chunk - 1) / chunk;
  if (nchunks > h->n_counters) {
    nchunks = h->n_counters;
    chunk = (N + nchunks - 1) / nchunks;
  }
  while ((int)h->chunk_ev.size() < nchunks) {
    cudaEvent_t e;
    OSC_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->chunk_ev.push_back(e);
  }
  // the copy stream may not overwrite inputs still being read by earlier work on `st`
  OSC_CUDA(h, cudaEventRecord(h->fence_ev, st));
  OSC_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->fence_ev, 0));
  const size_t B = sizeof(double);
  const size_t nv = h->nv, s = h->s, nc = h->nc, nu = h->nu;
  for (int c = 0; c < nchunks; ++c) {
    const size_t e0 = (size_t)c * chunk;
    const size_t n = (size_t)((e0 + chunk <= (size_t)N) ? chunk : N - e0);
    cudaStream_t cs = h->copy_stream;
    // task Jacobian: only the rows the kernels read (zero-weight task rows stay behind)
    if (J32) {
      const size_t F = sizeof(float);
      for (const auto& r : h->j_rows) {
        const size_t off = e0 * s * nv + (size_t)r.first * nv;
        OSC_CUDA(h, cudaMemcpy2DAsync(h->dJ32 + off, s * nv * F, J32 + off, s * nv * F,
                                      (size_t)(r.second - r.first) * nv * F, n,
                                      cudaMemcpyHostToDevice, cs));
      }
    } else if (h->j_rows.size() == 1 && h->j_rows[0].first == 0 && h->j_rows[0].second == (int)s) {
      OSC_CUDA(h, cudaMemcpyAsync(h->dJ + e0 * s * nv, J + e0 * s * nv, n * s * nv * B, cudaMemcpyHostToDevice, cs));
    } else {
      for (const auto& r : h->j_rows) {
        const size_t off = e0 * s * nv + (size_t)r.first * nv;
        OSC_CUDA(h, cudaMemcpy2DAsync(h->dJ + off, s * nv * B, J + off, s * nv * B,
                                      (size_t)(r.second - r.first) * nv * B, n,
                                      cudaMemcpyHostToDevice, cs));
      }
    }
    OSC_CUDA(h, cudaMemcpyAsync(h->dM + e0 * nv * nv, M + e0 * nv * nv, n * nv * nv * B, cudaMemcpyHostToDevice, cs));
    OSC_CUDA(h, cudaMemcpyAsync(h->dBias + e0 * s, bias + e0 * s, n * s * B, cudaMemcpyHostToDevice, cs));
    OSC_CUDA(h, cudaMemcpyAsync(h->dTargets + e0 * s, targets + e0 * s, n * s * B, cudaMemcpyHostToDevice, cs));
    OSC_CUDA(h, cudaMemcpyAsync(h->dC + e0 * nv, C + e0 * nv, n * nv * B, cudaMemcpyHostToDevice, cs));
    OSC_CUDA(h, cudaMemcpyAsync(h->dMask + e0 * nc, mask + e0 * nc, n * nc * B, cudaMemcpyHostToDevice, cs));
    OSC_CUDA(h, cudaEventRecord(h->chunk_ev[c], cs));
    OSC_CUDA(h, cudaStreamWaitEvent(st, h->chunk_ev[c], 0));
    if (J32) {  // widen on the compute stream: the copy stream keeps the link busy meanwhile
      for (const auto& r : h->j_rows) {
        const int len = (r.second - r.first) * (int)nv;
        int grid = (int)((n * len + 255) / 256);
        if (grid > h->sm_count * 8) grid = h->sm_count * 8;
        osc::widen_rows_kernel<<<grid, 256, 0, st>>>(h->dJ32 + e0 * s * nv, h->dJ + e0 * s * nv,
                                                     (int)n, (int)(s * nv), r.first * (int)nv, len);
        OSC_CUDA(h, cudaGetLastError());
        h->launches++;
      }
    }
    int rc = h->fuse_build ? OSC_OK : OSC_DISPATCH(h, launch_build, h, st, (int)e0, (int)n);
    if (rc) return rc;
    h->fuse_now = h->fuse_build;
    rc = OSC_DISPATCH(h, launch_solve, h, st, (int)e0, (int)n, c);
    h->fuse_now = false;
    if (rc) return rc;
    OSC_CUDA(h, cudaMemcpyAsync(torque + e0 * nu, h->dTorque + e0 * nu, n * nu * B, cudaMemcpyDeviceToHost, st));
  }
  h->kernels_ready = true;
  {
    size_t jrows = 0;
    for (const auto& r : h->j_rows) jrows += (size_t)(r.second - r.first);
    h->host_h2d_bytes = (size_t)N * B * (nv * nv + 2 * s + nv + nc) +
                        (size_t)N * jrows * nv * (J32 ? sizeof(float) : B);
    h->host_d2h_bytes = (size_t)N * B * nu;
  }
  OSC_CUDA(h, cudaStreamSynchronize(st));
  return OSC_OK;
}

int osc_step_host(osc_handle* h, const double* M, const double* C, const double* J,
                  const double* bias, const double* targets, const double* mask, double* torque,
                  void* stream) {
  return step_host_impl(h, M, C, J, nullptr, bias, targets, mask, torque, stream);
}

int osc_step_host_j32(osc_handle* h, const double* M, const double* C, const float* J32,
                      const double* bias, const double* targets, const double* mask,
                      double* torque, void* stream) {
  if (!J32) {
    if (h) h->err = "osc_step_host_j32: null host buffer";
    return OSC_ERR_INVALID;
  }
  return step_host_impl(h, M, C, nullptr, J32, bias, targets, mask, torque, stream);
}

#ifdef OSC_PHASE_CLOCKS
int osc_debug_phase_clocks(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, ::g_phase_clocks, 64 * sizeof(unsigned long long)) != cudaSuccess)
    return OSC_ERR_CUDA;
  if (reset) {
    unsigned long long z[64] = {0};
    if (cudaMemcpyToSymbol(::g_phase_clocks, z, sizeof(z)) != cudaSuccess) return OSC_ERR_CUDA;
  }
  return OSC_OK;
}
#endif

int osc_selftest_warp(int device, const double* in, double* out) {
  if (!in || !out) return OSC_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return OSC_ERR_CUDA;
  double *din = nullptr, *dout = nullptr;
  if (cudaMalloc((void**)&din, 18 * 32 * sizeof(double)) != cudaSuccess) return OSC_ERR_ALLOC;
  if (cudaMalloc((void**)&dout, 154 * sizeof(double)) != cudaSuccess) {
    cudaFree(din);
    return OSC_ERR_ALLOC;
  }
  cudaMemcpy(din, in, 18 * 32 * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 154 * sizeof(double));
  osc::warp_selftest_kernel<<<1, 32>>>(din, dout);
  const cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out, dout, 154 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(din);
  cudaFree(dout);
  return e == cudaSuccess ? OSC_OK : OSC_ERR_CUDA;
}

int osc_host_traffic(const osc_handle* h, size_t* h2d_bytes, size_t* d2h_bytes) {
  if (!h) return OSC_ERR_INVALID;
  if (h2d_bytes) *h2d_bytes = h->host_h2d_bytes;
  if (d2h_bytes) *d2h_bytes = h->host_d2h_bytes;
  return OSC_OK;
}

int osc_targets_pd(osc_handle* h, const osc_site_state* s, const double* kp_lin,
                   const double* kd_lin, const double* kp_ang, const double* kd_ang,
                   void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!s || !kp_lin || !kd_lin || !kp_ang || !kd_ang || !s->pos || !s->quat || !s->vel ||
      !s->angvel || !s->pos_des || !s->quat_des) {
    h->err = "osc_targets_pd: null argument";
    return OSC_ERR_INVALID;
  }
  if (h->iTargets != h->dTargets) {
    h->err = "osc_targets_pd: the targets input is bound to caller-owned memory";
    return OSC_ERR_STATE;
  }
  OSC_CUDA(h, cudaSetDevice(h->device));
  osc::PdGains g{};
  for (int i = 0; i < h->ns; ++i) {
    g.kp_lin[i] = kp_lin[i]; g.kd_lin[i] = kd_lin[i];
    g.kp_ang[i] = kp_ang[i]; g.kd_ang[i] = kd_ang[i];
  }
  const long long total = (long long)h->n_envs * h->ns;
  int grid = (int)((total + 255) / 256);
  if (grid > h->sm_count * 16) grid = h->sm_count * 16;
  osc::targets_pd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      g, s->pos, s->quat, s->vel, s->angvel, s->pos_des, s->quat_des, s->vel_des, s->angvel_des,
      h->dTargets, h->ns, total);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

int osc_walter_tumbling_default_gains(osc_walter_tumbling_gains* g) {
  if (!g) return OSC_ERR_INVALID;
  g->shin_kp = 800.0 * 3.0; g->shin_kv = 800.0 * 3.0; g->shin_rate = 0.1 * 8.0 * 5.0;
  g->thigh_kp = 4000.0 * 0.5; g->thigh_kv = 600.0 * 0.5; g->thigh_rate = 0.0;
  g->thigh_height_offset = -0.025;
  return OSC_OK;
}

int osc_targets_walter_tumbling(osc_handle* h, const osc_walter_tumbling_gains* g,
                                const double* shin_angle, const double* shin_angle_prev,
                                const double* shin_angle0, const double* thigh_z,
                                const double* thigh_z_prev, const double* thigh_z0, double time,
                                double dt, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!shin_angle || !shin_angle_prev || !shin_angle0 || !thigh_z || !thigh_z_prev || !thigh_z0 ||
      !(dt > 0.0)) {
    h->err = "osc_targets_walter_tumbling: null argument or dt <= 0";
    return OSC_ERR_INVALID;
  }
  if (h->ns != 17) {
    h->err = "osc_targets_walter_tumbling: the Walter site list (17 sites) only";
    return OSC_ERR_INVALID;
  }
  if (h->iTargets != h->dTargets) {
    h->err = "osc_targets_walter_tumbling: the targets input is bound to caller-owned memory";
    return OSC_ERR_STATE;
  }
  osc_walter_tumbling_gains gg;
  if (g) gg = *g; else osc_walter_tumbling_default_gains(&gg);
  OSC_CUDA(h, cudaSetDevice(h->device));
  const long long total = (long long)h->n_envs * h->ns;
  int grid = (int)((total + 255) / 256);
  if (grid > h->sm_count * 16) grid = h->sm_count * 16;
  osc::targets_walter_tumbling_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      gg, shin_angle, shin_angle_prev, shin_angle0, thigh_z, thigh_z_prev, thigh_z0, time, 1.0 / dt,
      h->dTargets, h->ns, total);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

int osc_contact_mask_from_contacts(osc_handle* h, const int* geom_pairs, const int* ncon,
                                   int max_con, const int* contact_geom_ids,
                                   const int* site_of_geom, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!geom_pairs || !ncon || !contact_geom_ids || max_con < 0) {
    h->err = "osc_contact_mask_from_contacts: bad argument";
    return OSC_ERR_INVALID;
  }
  if (h->iMask != h->dMask) {
    h->err = "osc_contact_mask_from_contacts: the mask input is bound to caller-owned memory";
    return OSC_ERR_STATE;
  }
  OSC_CUDA(h, cudaSetDevice(h->device));
  osc::ContactIds ids{};
  ids.nc = h->nc;
  for (int j = 0; j < h->nc; ++j) {
    ids.id[j] = contact_geom_ids[j];
    // getBinaryRepresentation_std_find(sites hit, list): entry c is raised when the site of
    // a touched listed geom equals list[c]
    const int site = site_of_geom ? site_of_geom[j] : contact_geom_ids[j];
    for (int c = 0; c < h->nc; ++c)
      if (contact_geom_ids[c] == site) ids.bits[j] |= 1u << c;
  }
  int grid = (h->n_envs + 255) / 256;
  osc::contact_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ids, geom_pairs, ncon, max_con,
                                                                  h->dMask, h->n_envs);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

// ---- multi-GPU gather of torques + statistics by peer stores --------------------------------
static size_t gather_slab_doubles(const osc_handle* h, int world) {
  return (size_t)world * h->n_envs * h->nu + (size_t)world * osc::kGatherStats;
}

int osc_gather_create(osc_handle* h, int rank, int world, void* ipc_handle_out) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (world < 1 || world > osc::kMaxPeers || rank < 0 || rank >= world) {
    h->err = "osc_gather_create: rank / world out of range (world <= 16)";
    return OSC_ERR_INVALID;
  }
  if (h->g_slab) {
    h->err = "osc_gather_create: already created";
    return OSC_ERR_STATE;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == OSC_IPC_HANDLE_BYTES, "IPC handle size");
  OSC_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = gather_slab_doubles(h, world) * sizeof(double);
  OSC_CUDA(h, cudaMalloc((void**)&h->g_slab, bytes));
  OSC_CUDA(h, cudaMemset(h->g_slab, 0, bytes));
  h->g_rank = rank;
  h->g_world = world;
  h->g_peer[rank] = h->g_slab;
  h->g_attached = world == 1;
  if (ipc_handle_out) {
    cudaIpcMemHandle_t ih;
    OSC_CUDA(h, cudaIpcGetMemHandle(&ih, h->g_slab));
    std::memcpy(ipc_handle_out, &ih, sizeof(ih));
  }
  return OSC_OK;
}

int osc_gather_attach(osc_handle* h, const void* ipc_handles, double* const* peer_slabs) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!h->g_slab) {
    h->err = "osc_gather_attach: osc_gather_create has not been called";
    return OSC_ERR_STATE;
  }
  if (!ipc_handles && !peer_slabs && h->g_world > 1) {
    h->err = "osc_gather_attach: neither IPC handles nor peer pointers given";
    return OSC_ERR_INVALID;
  }
  OSC_CUDA(h, cudaSetDevice(h->device));
  for (int p = 0; p < h->g_world; ++p) {
    if (p == h->g_rank) continue;
    if (peer_slabs) {  // same process: the peers' slabs are ordinary device pointers
      if (!peer_slabs[p]) {
        h->err = "osc_gather_attach: null peer pointer";
        return OSC_ERR_INVALID;
      }
      cudaPointerAttributes pa;
      OSC_CUDA(h, cudaPointerGetAttributes(&pa, peer_slabs[p]));
      if (pa.device != h->device) {
        int can = 0;
        OSC_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, pa.device));
        if (!can) {
          h->err = "osc_gather_attach: no peer access between the two devices";
          return OSC_ERR_CUDA;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(pa.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          h->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
          return OSC_ERR_CUDA;
        }
        (void)cudaGetLastError();
      }
      h->g_peer[p] = peer_slabs[p];
    } else {  // another process: map its slab (peer access is enabled by the open call)
      cudaIpcMemHandle_t ih;
      std::memcpy(&ih, static_cast<const unsigned char*>(ipc_handles) + (size_t)p * sizeof(ih),
                  sizeof(ih));
      void* ptr = nullptr;
      OSC_CUDA(h, cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
      h->g_peer[p] = static_cast<double*>(ptr);
      h->g_ipc_opened[p] = true;
    }
  }
  h->g_attached = true;
  return OSC_OK;
}

int osc_gather_torques(osc_handle* h, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!h->g_slab || !h->g_attached) {
    h->err = "osc_gather_torques: osc_gather_create / osc_gather_attach first";
    return OSC_ERR_STATE;
  }
  OSC_CUDA(h, cudaSetDevice(h->device));
  osc::GatherArgs a{};
  for (int p = 0; p < h->g_world; ++p) a.peer[p] = h->g_peer[p];
  a.torque = h->dTorque; a.iters = h->dIters; a.status = h->dStatus;
  a.pri_res = h->dPri; a.dua_res = h->dDua; a.reinits = h->dCounter + h->n_counters;
  a.slice = (size_t)h->n_envs * h->nu;
  a.stats_off = (size_t)h->g_world * a.slice;
  a.rank = h->g_rank; a.world = h->g_world; a.n_envs = h->n_envs;
  a.step = (double)(++h->g_steps);
  const size_t pairs = a.slice / 2;
  int grid = (int)((pairs + osc::kGatherThreads - 1) / osc::kGatherThreads);
  if (grid > h->sm_count) grid = h->sm_count;
  osc::gather_push_kernel<<<grid + 1, osc::kGatherThreads, 0, (cudaStream_t)stream>>>(a);
  OSC_CUDA(h, cudaGetLastError());
  h->launches++;
  return OSC_OK;
}

int osc_gather_buffers(osc_handle* h, double** torque_all, double** stats_all) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!h->g_slab) {
    h->err = "osc_gather_buffers: osc_gather_create has not been called";
    return OSC_ERR_STATE;
  }
  if (torque_all) *torque_all = h->g_slab;
  if (stats_all) *stats_all = h->g_slab + (size_t)h->g_world * h->n_envs * h->nu;
  return OSC_OK;
}

int osc_kinematics(osc_handle* h, const osc_kin_model* model, const double* qpos,
                   const double* qvel, void* stream) {
  if (check_handle(h)) return OSC_ERR_INVALID;
  if (!model || !qpos || !qvel) {
    h->err = "osc_kinematics: null argument";
    return OSC_ERR_INVALID;
  }
  if (model->nb < 1 || model->nb > OSC_KIN_MAX_BODIES || 6 + model->nb - 1 != h->nv ||
      model->ns != h->ns || model->parent[0] != -1) {
    h->err = "osc_kinematics: model does not match the robot shape (nv = 6 + nb - 1, ns)";
    return OSC_ERR_INVALID;
  }
  for (int b = 1; b < model->nb; ++b)
    if (model->parent[b] < 0 || model->parent[b] >= b) {
      h->err = "osc_kinematics: bodies must be in topological order (parent[b] < b)";
      return OSC_ERR_INVALID;
    }
  for (int s = 0; s < model->ns; ++s)
    if (model->site_body[s] < 0 || model->site_body[s] >= model->nb) {
      h->err = "osc_kinematics: site_body out of range";
      return OSC_ERR_INVALID;
    }
  if (h->iM != h->dM || h->iC != h->dC || h->iJ != h->dJ || h->iBias != h->dBias) {
    h->err = "osc_kinematics: M, C, J or bias is bound to caller-owned memory";
    return OSC_ERR_STATE;
  }
  OSC_CUDA(h, cudaSetDevice(h->device));
  return h->shape == osc::Shape::kWalter
             ? launch_kinematics<osc::WalterDims>(h, model, qpos, qvel, (cudaStream_t)stream)
             : launch_kinematics<osc::Go2Dims>(h, model, qpos, qvel, (cudaStream_t)stream);
}

long long osc_kernel_launches(const osc_handle* h) { return h ? h->launches : 0; }

int osc_reinit_count(osc_handle* h, int* count, void* stream) {
  if (check_handle(h) || !count) return OSC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(h, cudaSetDevice(h->device));
  OSC_CUDA(h, cudaMemcpyAsync(count, h->dCounter + h->n_counters, sizeof(int), cudaMemcpyDeviceToHost, st));
  OSC_CUDA(h, cudaStreamSynchronize(st));
  return OSC_OK;
}

int osc_measure_dfma_tflops(int device, double* tflops) {
  if (!tflops) return OSC_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return OSC_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OSC_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * blocks * threads) != cudaSuccess) return OSC_ERR_ALLOC;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    osc::dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return OSC_ERR_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return OSC_OK;
}

}  // extern "C"
