// Scalar-vector and IPA kernels.
//
// Replaces the serial loops of the reference:
//   src/inner_product_proof.rs:83-84,139-156,171-172,216-225  (inner products, a/b fold, G/H fold)
//   src/inner_product_proof.rs:302-311                        (s-vector)
//   src/r1cs/prover.rs:674-701,746-756,781-789                (l(x), r(x), t-poly, factors)
//   src/r1cs/verifier.rs:473-514                              (verification scalars)
//   src/util.rs:35-110                                        (exp_iter, VecPoly3, Poly6)
// Scalars are Fr elements in Montgomery form (8 x u32), vectors are plain arrays in HBM.
#pragma once
#include "ctx.cuh"

namespace bp {

struct PowTable { fe p[32]; };   // p[k] = base^(2^k)

template <class F>
__device__ __forceinline__ fe pow_from_table(const PowTable& t, uint32_t e) {
    fe r = F::one();
    for (int k = 0; k < 32 && (e >> k); k++)
        if ((e >> k) & 1u) r = F::mul(r, t.p[k]);
    return r;
}

// out[i] = base^i
template <class C>
__global__ void __launch_bounds__(256) vec_pow_kernel(const __grid_constant__ PowTable t, fe* __restrict__ out, size_t n) {
    using F = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fe(out + i, pow_from_table<F>(t, (uint32_t)i));
}

// block-wide sum of NV field elements per thread -> thread 0 holds the totals
template <class F, int NV, int BLOCK>
__device__ __forceinline__ void block_sum(fe (&v)[NV], fe* smem /* NV * BLOCK */) {
    for (int k = 0; k < NV; k++) smem[k * BLOCK + threadIdx.x] = v[k];
    __syncthreads();
    for (int stride = BLOCK / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride)
            for (int k = 0; k < NV; k++)
                smem[k * BLOCK + threadIdx.x] = F::add(smem[k * BLOCK + threadIdx.x], smem[k * BLOCK + threadIdx.x + stride]);
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < NV; k++) v[k] = smem[k * BLOCK];
}

// second stage: sums `nparts` rows of NV partials -> out[NV]
template <class C, int NV>
__global__ void __launch_bounds__(128) vec_reduce_partials_kernel(const fe* __restrict__ parts, int nparts, fe* __restrict__ out) {
    using F = Fp<typename C::Fr>;
    __shared__ fe sm[NV * 128];
    fe v[NV];
    for (int k = 0; k < NV; k++) v[k] = F::zero();
    for (int r = threadIdx.x; r < nparts; r += 128)
        for (int k = 0; k < NV; k++) v[k] = F::add(v[k], ld_fe_rw(parts + (size_t)r * NV + k));
    block_sum<F, NV, 128>(v, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < NV; k++) st_fe(out + k, v[k]);
}

// ---- IPA round preparation (inner_product_proof.rs:83-122 / 171-200) -----------------------------
// With h = n/2:  sLG[i] = a[i]*gR_i, sLH[i] = b[h+i]*hL_i, sRG[i] = a[h+i]*gL_i, sRH[i] = b[i]*hR_i,
// where (gL,gR,hL,hR) = fG*Gf[i], fG*Gf[h+i], fH*Hf[i], fH*Hf[h+i] while per-element factors apply (first round;
// every round when the factor vectors are geometric, see ipa_create) and the uniform deferred factors fG, fH
// otherwise; block partials of c_L = <a_L,b_R>, c_R = <a_R,b_L>.
// Multi-GPU shards: a rank owns the generators with global index i = j*P + g (j = local index);
// the scalar vectors a, b, Gf, Hf are replicated and read at the global index, outputs are written at the local one.
struct ShardIdx { uint32_t P, g; };

template <class C>
__global__ void __launch_bounds__(128) ipa_prep_kernel(const fe* __restrict__ a, const fe* __restrict__ b, size_t h, ShardIdx sh,
                                                       const fe* __restrict__ Gf, const fe* __restrict__ Hf, fe fG, fe fH,
                                                       fe* __restrict__ sLG, fe* __restrict__ sLH, fe* __restrict__ sRG,
                                                       fe* __restrict__ sRH, fe* __restrict__ parts) {
    using F = Fp<typename C::Fr>;
    __shared__ fe sm[2 * 128];
    fe acc[2] = {F::zero(), F::zero()};
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j * sh.P + sh.g < h; j += (size_t)gridDim.x * blockDim.x) {
        const size_t i = j * sh.P + sh.g;
        fe aL = ld_fe_rw(a + i), aR = ld_fe_rw(a + h + i), bL = ld_fe_rw(b + i), bR = ld_fe_rw(b + h + i);
        fe gL = fG, gR = fG, hL = fH, hR = fH;
        if (Gf) {
            gL = F::mul(fG, ld_fe_rw(Gf + i)); gR = F::mul(fG, ld_fe_rw(Gf + h + i));
            hL = F::mul(fH, ld_fe_rw(Hf + i)); hR = F::mul(fH, ld_fe_rw(Hf + h + i));
        }
        st_fe(sLG + j, F::mul(aL, gR));
        st_fe(sLH + j, F::mul(bR, hL));
        st_fe(sRG + j, F::mul(aR, gL));
        st_fe(sRH + j, F::mul(bL, hR));
        acc[0] = F::add(acc[0], F::mul(aL, bR));
        acc[1] = F::add(acc[1], F::mul(aR, bL));
    }
    block_sum<F, 2, 128>(acc, sm);
    if (threadIdx.x == 0) { st_fe(parts + 2 * blockIdx.x, acc[0]); st_fe(parts + 2 * blockIdx.x + 1, acc[1]); }
}

// a[i] = a[i]*u + uinv*a[h+i] ; b[i] = b[i]*uinv + u*b[h+i]      (inner_product_proof.rs:140-141)
template <class C>
__global__ void __launch_bounds__(256) ipa_fold_scalars_kernel(fe* __restrict__ a, fe* __restrict__ b, size_t h, fe u, fe uinv) {
    using F = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    fe aL = ld_fe_rw(a + i), aR = ld_fe_rw(a + h + i), bL = ld_fe_rw(b + i), bR = ld_fe_rw(b + h + i);
    st_fe(a + i, F::add(F::mul(aL, u), F::mul(uinv, aR)));
    st_fe(b + i, F::add(F::mul(bL, uinv), F::mul(u, bR)));
}

// ---- late IPA rounds without generator folding ---------------------------------------------------
// Only L_j, R_j, a, b are observable (hard part 3), so once the vectors are short the
// generators are no longer folded (a 256-step dependent scalar multiplication per round, latency
// bound); instead round j's L and R are MSMs over the generators of the last folded stage s
// (n_s points) with expanded scalars:  G^(j-1)_i = sum_tau cG(tau) * G^(s)_{i + n_{j-1}*tau},
// cG(tau) = prod_m u_{j-1-m}^(+1 if bit m of tau else -1), and the mirror image for H.
// Each stage point belongs to exactly one of L_j / R_j; the other MSM gets a zero scalar (skipped).
struct NoFoldParams {
    fe u[32], uinv[32];     // challenges of the unfolded rounds, oldest first (round s+1 at index 0)
    int nu;                 // how many (= j-1-s)
    fe fG, fH;              // uniform deferred factors of stage s (times Gf[t], Hf[t] when those are given)
};

template <class C>
__global__ void __launch_bounds__(128) ipa_nofold_scalars_kernel(const fe* __restrict__ a, const fe* __restrict__ b, size_t ns, size_t ncur,
                                                                 ShardIdx sh, const fe* __restrict__ Gf, const fe* __restrict__ Hf,
                                                                 const __grid_constant__ NoFoldParams p, fe* __restrict__ sLG,
                                                                 fe* __restrict__ sLH, fe* __restrict__ sRG, fe* __restrict__ sRH) {
    using F = Fp<typename C::Fr>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // local stage index (ns = local stage length)
    if (t >= ns) return;
    const size_t h = ncur / 2;
    const size_t gi = t * sh.P + sh.g;   // global stage index
    size_t ip = gi & (ncur - 1);         // index at the current level (ncur = current global length)
    size_t tau = gi / ncur;
    fe cG = Gf ? F::mul(p.fG, ld_fe_rw(Gf + gi)) : p.fG;
    fe cH = Hf ? F::mul(p.fH, ld_fe_rw(Hf + gi)) : p.fH;
    // bit m of tau belongs to round (j-1-m): index nu-1-m in the oldest-first arrays
    for (int m = 0; m < p.nu; m++) {
        bool bit = (tau >> m) & 1u;
        cG = F::mul(cG, bit ? p.u[p.nu - 1 - m] : p.uinv[p.nu - 1 - m]);
        cH = F::mul(cH, bit ? p.uinv[p.nu - 1 - m] : p.u[p.nu - 1 - m]);
    }
    fe z = F::zero();
    if (ip >= h) {
        // G_R half -> L (scalar a_L), H_R half -> R (scalar b_L)
        st_fe(sLG + t, F::mul(ld_fe_rw(a + (ip - h)), cG));
        st_fe(sRG + t, z);
        st_fe(sRH + t, F::mul(ld_fe_rw(b + (ip - h)), cH));
        st_fe(sLH + t, z);
    } else {
        // G_L half -> R (scalar a_R), H_L half -> L (scalar b_R)
        st_fe(sRG + t, F::mul(ld_fe_rw(a + (h + ip)), cG));
        st_fe(sLG + t, z);
        st_fe(sLH + t, F::mul(ld_fe_rw(b + (h + ip)), cH));
        st_fe(sRH + t, z);
    }
}

// block partials of c_L = <a_L, b_R>, c_R = <a_R, b_L> only
template <class C>
__global__ void __launch_bounds__(128) ipa_cross_kernel(const fe* __restrict__ a, const fe* __restrict__ b, size_t h, ShardIdx sh,
                                                        fe* __restrict__ parts) {
    using F = Fp<typename C::Fr>;
    __shared__ fe sm[2 * 128];
    fe acc[2] = {F::zero(), F::zero()};
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j * sh.P + sh.g < h; j += (size_t)gridDim.x * blockDim.x) {
        const size_t i = j * sh.P + sh.g;
        fe aL = ld_fe_rw(a + i), aR = ld_fe_rw(a + h + i), bL = ld_fe_rw(b + i), bR = ld_fe_rw(b + h + i);
        acc[0] = F::add(acc[0], F::mul(aL, bR));
        acc[1] = F::add(acc[1], F::mul(aR, bL));
    }
    block_sum<F, 2, 128>(acc, sm);
    if (threadIdx.x == 0) { st_fe(parts + 2 * blockIdx.x, acc[0]); st_fe(parts + 2 * blockIdx.x + 1, acc[1]); }
}

struct ScalarBits { uint32_t w[8]; };   // canonical (non-Montgomery) little-endian limbs

// Generator fold with one scalar shared by every thread (uniform control flow):
//   out[i] = L[i] + kappa * R[i]   -> affine.
// The deferred-factor form of inner_product_proof.rs:219-224: the reference computes
// u^-1*G_L + u*G_R = u^-1 * (G_L + u^2*G_R); the common factor is carried as a scalar
// (fG, fH) into the MSM scalars of the next round, so only one scalar multiplication per output
// point remains. Threads [0,count) fold (L0,R0) with kappa0, threads [count,2*count) fold (L1,R1)
// with kappa1 (G and H in one launch).
// XYZZ / extended -> affine for a whole 128-thread block with ONE field inversion (Montgomery's trick across the block).
// A fold output costs ~2 250 modmul of double-and-add and, with a private Fermat inversion, another ~450: one sixth of the
// kernel. Here every thread contributes z_i = ZZ*ZZZ (Z on the twisted Edwards curve), the warps build
// prefix and suffix products with shuffles (5 + 5 multiplications per thread), warp 0 inverts the product of the four
// warp totals -- one inversion stream for 128 outputs; the other warps wait at the barrier and leave the multiplier to
// the other resident blocks -- and 1/z_i = 1/total * (product of everything but z_i). Threads without an output, and
// identity points, take part with z = 1. Must be reached by all 128 threads of the block.
__device__ __forceinline__ fe shfl_fe(const fe& v, int src_lane) {
    fe r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(0xFFFFFFFFu, v.v[k], src_lane);
    return r;
}
template <class E>
__device__ __forceinline__ affine block_to_affine_128(const xyzz& p, bool valid) {
    using F = typename E::F;
    __shared__ fe sh_tot[4];
    __shared__ fe sh_inv[4];
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
    const bool live = valid && !E::is_identity(p);
    const bool te = E::IS_TE;
    const fe z = live ? (te ? p.zz : F::mul(p.zz, p.zzz)) : F::one();
    fe pre = z, suf = z;                                  // inclusive prefix / suffix products inside the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const fe up = shfl_fe(pre, lane >= d ? lane - d : lane);
        const fe dn = shfl_fe(suf, lane + d < 32 ? lane + d : lane);
        if (lane >= d) pre = F::mul(pre, up);
        if (lane + d < 32) suf = F::mul(suf, dn);
    }
    if (lane == 31) sh_tot[warp] = pre;
    __syncthreads();
    if (warp == 0) {
        const fe t0 = sh_tot[0], t1 = sh_tot[1], t2 = sh_tot[2], t3 = sh_tot[3];
        const fe t01 = F::mul(t0, t1), t23 = F::mul(t2, t3);
        const fe inv = F::inv(F::mul(t01, t23));          // the one inversion of the block
        if (lane < 4) {
            // 1 / t_lane = inv * (product of the other three warp totals)
            const fe others = lane == 0 ? F::mul(t1, t23) : lane == 1 ? F::mul(t0, t23) : lane == 2 ? F::mul(t01, t3) : F::mul(t01, t2);
            sh_inv[lane] = F::mul(inv, others);
        }
    }
    __syncthreads();
    // everything in this warp but z_i: exclusive prefix (lane - 1) times exclusive suffix (lane + 1)
    const fe pe = shfl_fe(pre, lane > 0 ? lane - 1 : 0), se = shfl_fe(suf, lane < 31 ? lane + 1 : 31);
    fe w = sh_inv[warp];
    if (lane > 0) w = F::mul(w, pe);
    if (lane < 31) w = F::mul(w, se);
    affine r = E::affine_identity();
    if (live) {
        if (te) {
            r.x = F::mul(p.x, w);
            r.y = F::mul(p.y, w);
        } else {
            r.x = F::mul(p.x, F::mul(w, p.zzz));          // 1/ZZ = ZZZ / (ZZ*ZZZ)
            r.y = F::mul(p.y, F::mul(w, p.zz));
        }
    }
    return r;
}

template <class C>
__global__ void __launch_bounds__(128) ipa_fold_points_uniform_kernel(const affine* __restrict__ L0, const affine* __restrict__ R0,
                                                                      affine* __restrict__ out0, const affine* __restrict__ L1,
                                                                      const affine* __restrict__ R1, affine* __restrict__ out1,
                                                                      size_t count, const __grid_constant__ ScalarBits k0,
                                                                      const __grid_constant__ ScalarBits k1, const __grid_constant__ ScalarBits k0x,
                                                                      const __grid_constant__ ScalarBits k1x, size_t cross_lo, size_t cross_hi, ShardIdx sh) {
    using E = GroupLaw<C>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < 2 * count;
    if (!valid) t = 0;                // keeps the block together for the shared inversion at the end
    const bool second = t >= count;   // count is a multiple of the block size or the block is split; either way correct
    size_t i = second ? t - count : t;
    const affine* L = second ? L1 : L0;
    const affine* R = second ? R1 : R0;
    affine* out = second ? out1 : out0;
    // piecewise-geometric factor vectors: elements whose partner lies across the block boundary use the second scalar
    const size_t gi = i * sh.P + sh.g;
    const bool cross = gi >= cross_lo && gi < cross_hi;
    const ScalarBits& k = second ? (cross ? k1x : k1) : (cross ? k0x : k0);
    affine pr = ld_affine(R + i);
    xyzz acc = E::identity();
    for (int limb = 7; limb >= 0; limb--) {
        uint32_t w = k.w[limb];
        for (int bit = 31; bit >= 0; bit--) {
            acc = E::dbl(acc);
            if ((w >> bit) & 1u) E::madd(acc, pr);
        }
    }
    affine pl = ld_affine(L + i);
    E::madd(acc, pl);
    affine r = block_to_affine_128<E>(acc, valid);
    if (!valid) return;
    st_fe(&out[i].x, r.x);
    st_fe(&out[i].y, r.y);
}

// The same fold on curves with the GLV endomorphism phi(x, y) = (beta*x, y) = lambda*(x, y) (secq256k1): the host splits
// kappa = k1 + k2*lambda with |k1|, |k2| < 2^130 (host/glv_host.hpp), so out[i] = L[i] + k1*R[i] + k2*phi(R[i]) is a joint
// double-and-add of ~129 steps over the table {+-R, +-phi(R), their sum} instead of 256 steps. The split is uniform
// over the launch, so control flow stays uniform.
struct GlvBits { uint32_t k1[5], k2[5]; int neg1, neg2, top; };

template <class C>
__global__ void __launch_bounds__(128) ipa_fold_points_glv_kernel(const affine* __restrict__ L0, const affine* __restrict__ R0,
                                                                  affine* __restrict__ out0, const affine* __restrict__ L1,
                                                                  const affine* __restrict__ R1, affine* __restrict__ out1,
                                                                  size_t count, const __grid_constant__ GlvBits g0,
                                                                  const __grid_constant__ GlvBits g1, const __grid_constant__ GlvBits g0x,
                                                                  const __grid_constant__ GlvBits g1x, size_t cross_lo, size_t cross_hi, ShardIdx sh) {
    using E = GroupLaw<C>;
    using F = Fp<typename C::Fq>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < 2 * count;
    if (!valid) t = 0;
    const bool second = t >= count;
    size_t i = second ? t - count : t;
    const affine* L = second ? L1 : L0;
    const affine* R = second ? R1 : R0;
    affine* out = second ? out1 : out0;
    const size_t gi = i * sh.P + sh.g;
    const bool cross = gi >= cross_lo && gi < cross_hi;     // see ipa_fold_points_uniform_kernel
    const GlvBits& g = second ? (cross ? g1x : g1) : (cross ? g0x : g0);
    affine p1 = ld_affine(R + i);
    affine p2;
    fe beta;
#pragma unroll
    for (int k = 0; k < 8; k++) beta.v[k] = C::glv_beta(k);
    p2.x = F::mul(p1.x, beta);
    p2.y = p1.y;
    if (E::is_identity(p1)) p2 = p1;
    if (g.neg1) p1 = E::neg(p1);
    if (g.neg2) p2 = E::neg(p2);
    xyzz both = E::from_affine(p1);
    E::madd(both, p2);
    xyzz acc = E::identity();
    for (int bit = g.top; bit >= 0; bit--) {
        acc = E::dbl(acc);
        const uint32_t sel = ((g.k1[bit >> 5] >> (bit & 31)) & 1u) | (((g.k2[bit >> 5] >> (bit & 31)) & 1u) << 1);
        if (sel == 1) E::madd(acc, p1);
        else if (sel == 2) E::madd(acc, p2);
        else if (sel == 3) E::add(acc, both);
    }
    affine pl = ld_affine(L + i);
    E::madd(acc, pl);
    affine r = block_to_affine_128<E>(acc, valid);
    if (!valid) return;
    st_fe(&out[i].x, r.x);
    st_fe(&out[i].y, r.y);
}

// The GLV fold with joint-sparse-form digits (host/glv_host.hpp: jsf_digits): the chain adds one of
// {+-p1, +-p2, +-(p1 + p2), +-(p1 - p2)} in about half of its ~130 steps (the binary joint chain above adds in three
// quarters of them, a quarter of those as full XYZZ additions of the un-normalised p1 + p2). p1 + p2 and p1 - p2 are
// made affine for the whole block with one shared inversion, so every addition of the chain is a mixed one:
// ~130 * (dbl + madd / 2) instead of ~130 * (dbl + madd / 2 + add / 4) modmul per output. Uniform control flow as before.
struct JsfBits { uint32_t code[21]; int neg1, neg2, top; };

// two XYZZ points -> affine with one field inversion per 128-thread block (see block_to_affine_128)
template <class E>
__device__ __forceinline__ void block_to_affine_128_x2(const xyzz& p, const xyzz& q, bool valid, affine& rp, affine& rq) {
    using F = typename E::F;
    static_assert(!E::IS_TE, "short Weierstrass only");
    __shared__ fe sh_tot2[4];
    __shared__ fe sh_inv2[4];
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
    const bool lp = valid && !E::is_identity(p), lq = valid && !E::is_identity(q);
    const fe zp = lp ? F::mul(p.zz, p.zzz) : F::one();
    const fe zq = lq ? F::mul(q.zz, q.zzz) : F::one();
    const fe z = F::mul(zp, zq);
    fe pre = z, suf = z;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const fe up = shfl_fe(pre, lane >= d ? lane - d : lane);
        const fe dn = shfl_fe(suf, lane + d < 32 ? lane + d : lane);
        if (lane >= d) pre = F::mul(pre, up);
        if (lane + d < 32) suf = F::mul(suf, dn);
    }
    if (lane == 31) sh_tot2[warp] = pre;
    __syncthreads();
    if (warp == 0) {
        const fe t0 = sh_tot2[0], t1 = sh_tot2[1], t2 = sh_tot2[2], t3 = sh_tot2[3];
        const fe t01 = F::mul(t0, t1), t23 = F::mul(t2, t3);
        const fe inv = F::inv(F::mul(t01, t23));
        if (lane < 4) {
            const fe others = lane == 0 ? F::mul(t1, t23) : lane == 1 ? F::mul(t0, t23) : lane == 2 ? F::mul(t01, t3) : F::mul(t01, t2);
            sh_inv2[lane] = F::mul(inv, others);
        }
    }
    __syncthreads();
    const fe pe = shfl_fe(pre, lane > 0 ? lane - 1 : 0), se = shfl_fe(suf, lane < 31 ? lane + 1 : 31);
    fe w = sh_inv2[warp];
    if (lane > 0) w = F::mul(w, pe);
    if (lane < 31) w = F::mul(w, se);              // w = 1 / (zp * zq)
    rp = E::affine_identity();
    rq = E::affine_identity();
    if (lp) {
        const fe wp = F::mul(w, zq);               // 1 / zp
        rp.x = F::mul(p.x, F::mul(wp, p.zzz));
        rp.y = F::mul(p.y, F::mul(wp, p.zz));
    }
    if (lq) {
        const fe wq = F::mul(w, zp);
        rq.x = F::mul(q.x, F::mul(wq, q.zzz));
        rq.y = F::mul(q.y, F::mul(wq, q.zz));
    }
}

template <class C>
__global__ void __launch_bounds__(128) ipa_fold_points_jsf_kernel(const affine* __restrict__ L0, const affine* __restrict__ R0,
                                                                  affine* __restrict__ out0, const affine* __restrict__ L1,
                                                                  const affine* __restrict__ R1, affine* __restrict__ out1,
                                                                  size_t count, const __grid_constant__ JsfBits g0,
                                                                  const __grid_constant__ JsfBits g1, const __grid_constant__ JsfBits g0x,
                                                                  const __grid_constant__ JsfBits g1x, size_t cross_lo, size_t cross_hi, ShardIdx sh) {
    using E = GroupLaw<C>;
    using F = Fp<typename C::Fq>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < 2 * count;
    if (!valid) t = 0;
    const bool second = t >= count;
    size_t i = second ? t - count : t;
    const affine* L = second ? L1 : L0;
    const affine* R = second ? R1 : R0;
    affine* out = second ? out1 : out0;
    const size_t gi = i * sh.P + sh.g;
    const bool cross = gi >= cross_lo && gi < cross_hi;     // see ipa_fold_points_uniform_kernel
    const JsfBits& g = second ? (cross ? g1x : g1) : (cross ? g0x : g0);
    affine p1 = ld_affine(R + i);
    affine p2;
    fe beta;
#pragma unroll
    for (int k = 0; k < 8; k++) beta.v[k] = C::glv_beta(k);
    p2.x = F::mul(p1.x, beta);
    p2.y = p1.y;
    if (E::is_identity(p1)) p2 = p1;
    if (g.neg1) p1 = E::neg(p1);
    if (g.neg2) p2 = E::neg(p2);
    affine ps, pd;
    {
        xyzz s = E::from_affine(p1), d = E::from_affine(p1);
        E::madd(s, p2);
        E::madd(d, E::neg(p2));
        block_to_affine_128_x2<E>(s, d, valid, ps, pd);
    }
    xyzz acc = E::identity();
#pragma unroll 1
    for (int step = g.top; step >= 0; step--) {
        acc = E::dbl(acc);
        const uint32_t code = (g.code[step >> 3] >> (4 * (step & 7))) & 15u;
        if (code == 5u) continue;                            // (0, 0)
        const int u1 = (int)(code & 3u) - 1, u2 = (int)(code >> 2) - 1;
        // u1*p1 + u2*p2 = +-p1, +-p2, +-(p1 + p2) or +-(p1 - p2)
        affine q;
        bool ng;
        if (u2 == 0) { q = p1; ng = u1 < 0; }
        else if (u1 == 0) { q = p2; ng = u2 < 0; }
        else if (u1 == u2) { q = ps; ng = u1 < 0; }
        else { q = pd; ng = u1 < 0; }
        if (ng) q = E::neg(q);
        E::madd(acc, q);
    }
    affine pl = ld_affine(L + i);
    E::madd(acc, pl);
    affine r = block_to_affine_128<E>(acc, valid);
    if (!valid) return;
    st_fe(&out[i].x, r.x);
    st_fe(&out[i].y, r.y);
}

// First-round fold with per-element factors (inner_product_proof.rs:143-155):
//   out[i] = (cL*f[i]) * P[i] + (cR*f[h+i]) * P[h+i]      (joint double-and-add)
// Threads [0,h) handle (P0,f0,cL0,cR0) = (G, G_factors, u^-1, u); threads [h,2h) handle H with
// (H_factors, u, u^-1).
template <class C>
__global__ void __launch_bounds__(128) ipa_fold_points_joint_kernel(const affine* __restrict__ P0, const fe* __restrict__ f0, fe cL0, fe cR0,
                                                                    affine* __restrict__ out0, const affine* __restrict__ P1,
                                                                    const fe* __restrict__ f1, fe cL1, fe cR1, affine* __restrict__ out1,
                                                                    size_t h, size_t hl, ShardIdx sh) {
    using E = GroupLaw<C>;
    using Fr = Fp<typename C::Fr>;
    // h = global half length (factor vectors are replicated), hl = local half length (points are this rank's shard)
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < 2 * hl;
    if (!valid) t = 0;
    const bool second = t >= hl;
    size_t i = second ? t - hl : t;
    const size_t gi = i * sh.P + sh.g;
    const affine* P = second ? P1 : P0;
    const fe* f = second ? f1 : f0;
    affine* out = second ? out1 : out0;
    fe sl = Fr::from_mont(Fr::mul(second ? cL1 : cL0, ld_fe_rw(f + gi)));
    fe sr = Fr::from_mont(Fr::mul(second ? cR1 : cR0, ld_fe_rw(f + h + gi)));
    affine pl = ld_affine(P + i), pr = ld_affine(P + hl + i);
    xyzz both = E::from_affine(pl);
    E::madd(both, pr);
    xyzz xl = E::from_affine(pl), xr = E::from_affine(pr);
    xyzz acc = E::identity();
    for (int limb = 7; limb >= 0; limb--) {
        uint32_t wl = sl.v[limb], wr = sr.v[limb];
        for (int bit = 31; bit >= 0; bit--) {
            acc = E::dbl(acc);
            uint32_t sel = ((wl >> bit) & 1u) | (((wr >> bit) & 1u) << 1);
            if (sel) {
                xyzz q = sel == 1 ? xl : sel == 2 ? xr : both;
                E::add(acc, q);
            }
        }
    }
    affine r = block_to_affine_128<E>(acc, valid);
    if (!valid) return;
    st_fe(&out[i].x, r.x);
    st_fe(&out[i].y, r.y);
}

// ---- batched Pedersen commitments (generators.rs:39-44 called m times: prover.rs:327-341) ------------
// out[i] = v[i]*B + r[i]*B_blinding. One thread per commitment, joint double-and-add over the shared
// table {B, B_blinding, B + B_blinding} (all affine, so every addition is a mixed add).
template <class C>
__global__ void __launch_bounds__(128) pedersen_commit_kernel(const affine B, const affine Bb, const affine BBb, const fe* __restrict__ v,
                                                              const fe* __restrict__ r, affine* __restrict__ out, size_t n) {
    using E = GroupLaw<C>;
    using Fr = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe sv = Fr::from_mont(ld_fe(v + i)), sr = Fr::from_mont(ld_fe(r + i));
    xyzz acc = E::identity();
    for (int limb = 7; limb >= 0; limb--) {
        uint32_t wv = sv.v[limb], wr = sr.v[limb];
        for (int bit = 31; bit >= 0; bit--) {
            acc = E::dbl(acc);
            uint32_t sel = ((wv >> bit) & 1u) | (((wr >> bit) & 1u) << 1);
            if (sel) E::madd(acc, sel == 1 ? B : sel == 2 ? Bb : BBb);
        }
    }
    affine a = E::to_affine(acc);
    st_fe(&out[i].x, a.x);
    st_fe(&out[i].y, a.y);
}

// Fixed-base form of the same commitments ("fixed-base -> precomputed tables"): B and B_blinding
// never change, so a table T[b][j][d] = d * 2^(8j) * P_b (2 bases x 32 byte-windows x 256 digits, 1 MB, built once per
// generator set) turns a commitment into at most 64 mixed additions and no doubling -- the 256-step chain above is
// pure latency (2.5 ms per launch however few commitments there are), this is ~0.3 ms.
// pedersen_window_kernel: W[b][j] = 2^(8j) * P_b (one thread per (b, j)); pedersen_table_kernel: the 256 multiples.
template <class C>
__global__ void __launch_bounds__(64) pedersen_window_kernel(const affine B, const affine Bb, affine* __restrict__ win) {
    using E = GroupLaw<C>;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 64) return;
    const int b = t >> 5, j = t & 31;
    xyzz acc = E::from_affine(b ? Bb : B);
    for (int k = 0; k < 8 * j; k++) acc = E::dbl(acc);
    affine a = E::to_affine(acc);
    st_fe(&win[t].x, a.x);
    st_fe(&win[t].y, a.y);
}
template <class C>
__global__ void __launch_bounds__(128) pedersen_table_kernel(const affine* __restrict__ win, affine* __restrict__ table) {
    using E = GroupLaw<C>;
    int t = blockIdx.x * blockDim.x + threadIdx.x;       // (b*32 + j)*256 + d
    if (t >= 2 * 32 * 256) return;
    const uint32_t d = (uint32_t)t & 255u;
    affine w = ld_affine(win + (t >> 8));
    xyzz acc = E::mul_u32(E::from_affine(w), d);          // d = 0 -> identity -> (0, 0)
    affine a = E::to_affine(acc);
    st_fe(&table[t].x, a.x);
    st_fe(&table[t].y, a.y);
}
template <class C>
__global__ void __launch_bounds__(128) pedersen_commit_table_kernel(const affine* __restrict__ table, const fe* __restrict__ v,
                                                                    const fe* __restrict__ r, affine* __restrict__ out, size_t n) {
    using E = GroupLaw<C>;
    using Fr = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe sv = Fr::from_mont(ld_fe(v + i)), sr = Fr::from_mont(ld_fe(r + i));
    xyzz acc = E::identity();
#pragma unroll 1
    for (int j = 0; j < 32; j++) {
        const uint32_t dv = (sv.v[j >> 2] >> (8 * (j & 3))) & 255u, dr = (sr.v[j >> 2] >> (8 * (j & 3))) & 255u;
        if (dv) { affine p = ld_affine(table + ((size_t)j << 8) + dv); E::madd(acc, p); }
        if (dr) { affine p = ld_affine(table + ((size_t)(32 + j) << 8) + dr); E::madd(acc, p); }
    }
    affine a = E::to_affine(acc);
    st_fe(&out[i].x, a.x);
    st_fe(&out[i].y, a.y);
}

// ---- R1CS prover vector kernels (prover.rs:674-756) ----------------------------------------------
struct LrInputs {
    const fe *aL, *aR, *aO, *sL, *sR, *wL, *wR, *wO;   // length n
    const fe *ypow, *yinvpow;                          // y^i, y^-i, length padded_n
};

// t-polynomial coefficients (util.rs:75-93): block partials of t1..t6
template <class C>
__global__ void __launch_bounds__(128) r1cs_tpoly_kernel(const __grid_constant__ LrInputs in, size_t n, fe* __restrict__ parts) {
    using F = Fp<typename C::Fr>;
    __shared__ fe sm[6 * 128];
    fe t[6];
    for (int k = 0; k < 6; k++) t[k] = F::zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fe yp = ld_fe_rw(in.ypow + i), yi = ld_fe_rw(in.yinvpow + i);
        fe l1 = F::add(ld_fe_rw(in.aL + i), F::mul(yi, ld_fe_rw(in.wR + i)));     // prover.rs:687
        fe l2 = ld_fe_rw(in.aO + i);                                               // :689
        fe l3 = ld_fe_rw(in.sL + i);                                               // :691
        fe r0 = F::sub(ld_fe_rw(in.wO + i), yp);                                   // :693
        fe r1 = F::add(F::mul(yp, ld_fe_rw(in.aR + i)), ld_fe_rw(in.wL + i));      // :695
        fe r3 = F::mul(yp, ld_fe_rw(in.sR + i));                                   // :698
        t[0] = F::add(t[0], F::mul(l1, r0));
        t[1] = F::add(t[1], F::add(F::mul(l1, r1), F::mul(l2, r0)));
        t[2] = F::add(t[2], F::add(F::mul(l2, r1), F::mul(l3, r0)));
        t[3] = F::add(t[3], F::add(F::mul(l1, r3), F::mul(l3, r1)));
        t[4] = F::add(t[4], F::mul(l2, r3));
        t[5] = F::add(t[5], F::mul(l3, r3));
    }
    block_sum<F, 6, 128>(t, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < 6; k++) st_fe(parts + 6 * blockIdx.x + k, t[k]);
}

// l_vec = l(x), r_vec = r(x) with padding (util.rs:95-102, prover.rs:746-756), and the IPA factor
// vectors G_factors = 1^{n1} || u^{..}, H_factors = y^-i * G_factors (prover.rs:781-789).
template <class C>
__global__ void __launch_bounds__(256) r1cs_lr_eval_kernel(const __grid_constant__ LrInputs in, size_t n, size_t padded_n, size_t n1,
                                                           fe x, fe u, fe* __restrict__ l_vec, fe* __restrict__ r_vec,
                                                           fe* __restrict__ Gf, fe* __restrict__ Hf) {
    using F = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded_n) return;
    fe yp = ld_fe_rw(in.ypow + i), yi = ld_fe_rw(in.yinvpow + i);
    fe l, r;
    if (i < n) {
        fe l1 = F::add(ld_fe_rw(in.aL + i), F::mul(yi, ld_fe_rw(in.wR + i)));
        fe l2 = ld_fe_rw(in.aO + i);
        fe l3 = ld_fe_rw(in.sL + i);
        fe r0 = F::sub(ld_fe_rw(in.wO + i), yp);
        fe r1 = F::add(F::mul(yp, ld_fe_rw(in.aR + i)), ld_fe_rw(in.wL + i));
        fe r3 = F::mul(yp, ld_fe_rw(in.sR + i));
        l = F::mul(x, F::add(l1, F::mul(x, F::add(l2, F::mul(x, l3)))));
        r = F::add(r0, F::mul(x, F::add(r1, F::mul(x, F::mul(x, r3)))));
    } else {
        l = F::zero();
        r = F::neg(yp);
    }
    st_fe(l_vec + i, l);
    st_fe(r_vec + i, r);
    fe g = i < n1 ? F::one() : u;
    st_fe(Gf + i, g);
    st_fe(Hf + i, F::mul(yi, g));
}

// ---- verifier scalars (verifier.rs:473-514, inner_product_proof.rs:302-311) -----------------------
struct VerifyInputs {
    const fe *wL, *wR, *wO;     // length n
    const fe* yinvpow;          // length padded_n
    fe usq[32];                 // u_j^2 in creation order (challenges_sq)
    fe allinv, x, a, b, u;
    int lg_n;
};

// s_i = allinv * prod_{j: bit j of i set} u_sq[lg_n-1-j]
template <class F>
__device__ __forceinline__ fe s_value(const VerifyInputs& in, uint32_t i) {
    fe s = in.allinv;
    for (int j = 0; j < in.lg_n; j++)
        if ((i >> j) & 1u) s = F::mul(s, in.usq[in.lg_n - 1 - j]);
    return s;
}

template <class C>
__global__ void __launch_bounds__(128) r1cs_verify_scalars_kernel(const __grid_constant__ VerifyInputs in, size_t n, size_t padded_n,
                                                                  size_t n1, fe* __restrict__ g_scalars, fe* __restrict__ h_scalars,
                                                                  fe* __restrict__ delta_parts) {
    using F = Fp<typename C::Fr>;
    __shared__ fe sm[128];
    fe acc[1] = {F::zero()};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded_n; i += (size_t)gridDim.x * blockDim.x) {
        fe yi = ld_fe_rw(in.yinvpow + i);
        fe wl = F::zero(), wr = F::zero(), wo = F::zero();
        if (i < n) { wl = ld_fe_rw(in.wL + i); wr = ld_fe_rw(in.wR + i); wo = ld_fe_rw(in.wO + i); }
        fe yneg_wr = F::mul(wr, yi);                                                         // verifier.rs:477-482
        acc[0] = F::add(acc[0], F::mul(yneg_wr, wl));                                         // delta, :484
        fe si = s_value<F>(in, (uint32_t)i);
        fe sinv = s_value<F>(in, (uint32_t)(padded_n - 1 - i));                               // 1/s_i = s_{n-1-i}
        fe g = F::sub(F::mul(in.x, yneg_wr), F::mul(in.a, si));                               // :496
        fe hh = F::sub(F::mul(yi, F::sub(F::add(F::mul(in.x, wl), wo), F::mul(in.b, sinv))), F::one());   // :512
        if (i >= n1) { g = F::mul(in.u, g); hh = F::mul(in.u, hh); }
        st_fe(g_scalars + i, g);
        st_fe(h_scalars + i, hh);
    }
    block_sum<F, 1, 128>(acc, sm);
    if (threadIdx.x == 0) st_fe(delta_parts + blockIdx.x, acc[0]);
}

// InnerProductProof::verify scalars (inner_product_proof.rs:338-353):
//   g[i] = (a * s_i) * G_factors[i] ,  h[i] = (b * s_{n-1-i}) * H_factors[i]
template <class C>
__global__ void __launch_bounds__(128) ipa_verify_scalars_kernel(const __grid_constant__ VerifyInputs in, size_t n, const fe* __restrict__ Gf,
                                                                 const fe* __restrict__ Hf, fe* __restrict__ g, fe* __restrict__ h) {
    using F = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe si = s_value<F>(in, (uint32_t)i);
    fe sinv = s_value<F>(in, (uint32_t)(n - 1 - i));
    st_fe(g + i, F::mul(F::mul(in.a, si), ld_fe_rw(Gf + i)));
    st_fe(h + i, F::mul(F::mul(in.b, sinv), ld_fe_rw(Hf + i)));
}

// out[i] = alpha * in[i]  (+ accumulate into acc[i] if acc != nullptr)      (verifier.rs:650-664)
template <class C>
__global__ void __launch_bounds__(256) vec_scale_accum_kernel(const fe* __restrict__ in, fe alpha, fe* __restrict__ acc, size_t n) {
    using F = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_fe(acc + i, F::add(ld_fe_rw(acc + i), F::mul(alpha, ld_fe_rw(in + i))));
}

}  // namespace bp
