// C++ host mirror of the reference's proving / verification API, driving the CUDA kernels.
//
//   InnerProductProof::create          src/inner_product_proof.rs:37-239     -> ipa_create()
//   Prover {commit, multiply, allocate, allocate_multiplier, constrain,
//           specify_randomized_constraints, prove}   src/r1cs/prover.rs:96-831 -> ProverT<C>
//   Verifier {commit, ..., verification_scalars, verify}  src/r1cs/verifier.rs:69-600 -> VerifierT<C>
//   batch_verify                        src/r1cs/verifier.rs:604-691          -> batch_verify_t()
//   R1CSProof::{to_bytes,from_bytes}    src/r1cs/proof.rs:25-91               -> ProofT<C>
//
// Same names, argument meaning and error behaviour as the Rust crate (Rust itself is absent from
// this image). The transcript schedule, RNG draw order and wire format are the reference's, so
// proofs are byte-identical to the CPU oracle; all O(n) arithmetic runs in the kernels of
// msm_kernels.cuh / vec_kernels.cuh on the context's stream.
#pragma once
#include <string.h>
#include <chrono>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include "host/gens_host.hpp"
#include "r1cs_types.hpp"
#include "msm_kernels.cuh"
#include "flatten.cuh"
#include "transcript_dev.cuh"
#include "host/glv_host.hpp"
#include "vec_kernels.cuh"

namespace bp {

template <class C>
struct ProofT : ProofBase {
    using HC = HostCurve<C>;
    affine A_I1, A_O1, S1, A_I2, A_O2, S2, T_1, T_3, T_4, T_5, T_6;
    fe t_x, t_x_blinding, e_blinding;
    std::vector<affine> L_vec, R_vec;
    fe a, b;

    std::vector<uint8_t> to_bytes() const override {                     // proof.rs:74-78
        std::vector<uint8_t> out;
        constexpr int PC = HC::POINT_COMPRESSED;
        auto pt = [&](const affine& p) { uint8_t buf[33] = {0}; HC::point_compressed(p, buf); out.insert(out.end(), buf, buf + PC); };
        auto sc = [&](const fe& s) { uint8_t buf[32]; HC::scalar_to_bytes(s, buf); out.insert(out.end(), buf, buf + 32); };
        auto u64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i))); };
        const affine* pts[11] = {&A_I1, &A_O1, &S1, &A_I2, &A_O2, &S2, &T_1, &T_3, &T_4, &T_5, &T_6};
        for (auto p : pts) pt(*p);
        sc(t_x); sc(t_x_blinding); sc(e_blinding);
        u64(L_vec.size()); for (auto& p : L_vec) pt(p);
        u64(R_vec.size()); for (auto& p : R_vec) pt(p);
        sc(a); sc(b);
        return out;
    }
    // `defer`: instead of decompressing each point on the host, record (compressed bytes, destination) pairs for one
    // batched GPU decompression over many proofs (api_impl.cuh: proofs_from_bytes_batch)
    static int from_bytes(const uint8_t* d, size_t len, ProofT& pr, std::vector<std::pair<const uint8_t*, affine*>>* defer = nullptr) {     // proof.rs:83-91 (FormatError)
        size_t off = 0;
        constexpr size_t PC = HC::POINT_COMPRESSED;
        auto pt = [&](affine& p) {
            if (off + PC > len) return false;
            bool ok = true;
            if (defer) defer->emplace_back(d + off, &p);
            else ok = HC::point_from_compressed(d + off, p);
            off += PC;
            return ok;
        };
        auto sc = [&](fe& s) { if (off + 32 > len) return false; bool ok = HC::scalar_from_bytes(d + off, s); off += 32; return ok; };
        affine* pts[11] = {&pr.A_I1, &pr.A_O1, &pr.S1, &pr.A_I2, &pr.A_O2, &pr.S2, &pr.T_1, &pr.T_3, &pr.T_4, &pr.T_5, &pr.T_6};
        for (auto p : pts) if (!pt(*p)) return BP_ERR_FORMAT;
        if (!sc(pr.t_x) || !sc(pr.t_x_blinding) || !sc(pr.e_blinding)) return BP_ERR_FORMAT;
        for (int v = 0; v < 2; v++) {
            if (off + 8 > len) return BP_ERR_FORMAT;
            uint64_t cnt = 0;
            for (int i = 0; i < 8; i++) cnt |= (uint64_t)d[off + i] << (8 * i);
            off += 8;
            if (cnt > (len - off) / PC) return BP_ERR_FORMAT;
            auto& vec = v == 0 ? pr.L_vec : pr.R_vec;
            vec.resize(cnt);
            for (auto& p : vec) if (!pt(p)) return BP_ERR_FORMAT;
        }
        if (!sc(pr.a) || !sc(pr.b)) return BP_ERR_FORMAT;
        return BP_OK;
    }
};

// ---- transcript protocol (src/transcript.rs:45-101) -------------------------------------------------
template <class C>
struct TP {
    using HC = HostCurve<C>;
    static void append_scalar(Transcript& t, const char* label, const fe& s) { uint8_t b[32]; HC::scalar_to_bytes(s, b); t.append_message(label, b, 32); }
    static void append_point(Transcript& t, const char* label, const affine& p) { uint8_t b[65] = {0}; HC::point_uncompressed(p, b); t.append_message(label, b, HC::POINT_UNCOMPRESSED); }
    static int validate_and_append_point(Transcript& t, const char* label, const affine& p) {
        if (HC::E::is_identity(p)) return BP_ERR_VERIFY;
        append_point(t, label, p);
        return BP_OK;
    }
    static fe challenge_scalar(Transcript& t, const char* label) {
        uint8_t buf[32];
        t.challenge_bytes(label, buf, 32);
        ChaCha20Rng rng(buf);
        return HC::scalar_rand(rng);
    }
};

struct StageTimer {
    bp_ctx* ctx;
    std::chrono::steady_clock::time_point t0;
    explicit StageTimer(bp_ctx* c) : ctx(c), t0(std::chrono::steady_clock::now()) { for (auto& x : c->stage_ms) x = 0; }
    void lap(int i) {
        auto t1 = std::chrono::steady_clock::now();
        ctx->stage_ms[i] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
    }
};
enum { ST_RNG = 0, ST_COMMIT = 1, ST_FLATTEN = 2, ST_VEC = 3, ST_TCOMMIT = 4, ST_IPA = 5, ST_IPA_MSM = 6, ST_IPA_FOLD = 7, ST_IPA_HOST = 8,
       ST_VSCALARS = 9, ST_VMSM = 10, ST_UPLOAD = 11, ST_TAIL = 12 };

// ---- device helpers -----------------------------------------------------------------------------
template <class C>
struct Dev {
    using Fr = HostFp<typename C::Fr>;
    static PowTable pow_table(const fe& base) {
        PowTable t;
        t.p[0] = base;
        for (int k = 1; k < 32; k++) t.p[k] = Fr::sqr(t.p[k - 1]);
        return t;
    }
    static int pow_vec(bp_ctx* ctx, const fe& base, fe* d_out, size_t n) {
        if (!n) return BP_OK;
        vec_pow_kernel<C><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(pow_table(base), d_out, n);
        BP_LAUNCH_CHECK(ctx);
        return BP_OK;
    }
    static int upload(bp_ctx* ctx, void* dst, const void* src, size_t bytes) {
        if (bytes) BP_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return BP_OK;
    }
    static int download(bp_ctx* ctx, void* dst, const void* src, size_t bytes) {
        if (bytes) BP_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return BP_OK;
    }
    static ScalarBits bits(const fe& mont) {
        fe c = Fr::from_mont(mont);
        ScalarBits b;
        for (int i = 0; i < 8; i++) b.w[i] = c.v[i];
        return b;
    }
};

// ---- InnerProductProof::create on the device (inner_product_proof.rs:37-239) ------------------------
// d_G, d_H: generators (read-only); d_Gf, d_Hf: factor vectors; d_a, d_b: overwritten.
template <class C>
int ipa_create(bp_ctx* ctx, Transcript& t, const affine& Q, const fe* d_Gf, const fe* d_Hf, const affine* d_G, const affine* d_H,
               fe* d_a, fe* d_b, size_t n, std::vector<affine>& L_vec, std::vector<affine>& R_vec, fe& a_out, fe& b_out,
               const fe* geo_rG = nullptr, const fe* geo_rH = nullptr, size_t geo_n1 = 0, const fe* geo_jump = nullptr) {
    using Fr = HostFp<typename C::Fr>;
    using D = Dev<C>;
    if (n == 0 || (n & (n - 1))) return BP_ERR_POW2;                     // assert at :66
    // Piecewise-geometric factor vectors: Gf[i+1] = rG*Gf[i], Hf[i+1] = rH*Hf[i], except for one jump by `geo_jump` at
    // index geo_n1. The R1CS prover's always are (prover.rs:781-789: G_f = [1]*n1 ++ [u]*(n2+pad), H_f[i] = y^-i*G_f[i]).
    // The two partners of a fold then differ by r^h, or by r^h*jump when the boundary lies between them, so
    //   cL*f[i]*P[i] + cR*f[h+i]*P[h+i] = cL*f[i]*(P[i] + (cR/cL)*r^h*[jump]*P[h+i]):
    // all rounds use the (two-)uniform-scalar fold, the generators stay unscaled, and the per-element factor
    // f[i] (a prefix of the original vector) times the running common factor goes into the MSM scalars.
    const bool geo = geo_rG && geo_rH;
    t.append_message("dom-sep", (const uint8_t*)"ipp v1", 6);            // transcript.rs:52-55
    t.append_u64("n", n);
    L_vec.clear();
    R_vec.clear();
    cudaStream_t st = ctx->stream;
    // Multi-GPU: d_G, d_H are this rank's cyclic shard (local j = global j*P + g), a, b and the
    // factor vectors are replicated. Index i and its fold partner i + n/2 share a rank while n >= 2P, so the
    // generator fold is local; each rank's L/R MSM runs over its shard (with its share of c_L, c_R on Q) and the
    // partial points are all-gathered and summed. Once n < 2P (or below the no-fold threshold) the no-fold form
    // is used, which needs no generator exchange at all.
    const size_t P = (size_t)ctx->world, g = (size_t)ctx->rank;
    if (P > 1 && (n < P || n % P)) return BP_ERR_ARG;
    size_t half = n / 2 / P;
    BP_CUDA_TRY(ctx, ctx->ipa_G.reserve((half + 1) * sizeof(affine)));
    BP_CUDA_TRY(ctx, ctx->ipa_H.reserve((half + 1) * sizeof(affine)));
    BP_CUDA_TRY(ctx, ctx->ipa_s.reserve((4 * half + 8) * sizeof(fe)));
    const int PREP_BLOCKS = 296;
    BP_CUDA_TRY(ctx, ctx->ipa_parts.reserve((size_t)(2 * PREP_BLOCKS + 8) * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->small.reserve(4096));
    affine* wG = ctx->ipa_G.as<affine>();
    affine* wH = ctx->ipa_H.as<affine>();
    fe* s_all = ctx->ipa_s.as<fe>();
    fe* parts = ctx->ipa_parts.as<fe>();
    fe* d_c = ctx->small.as<fe>();                      // [cL, cR]
    affine* d_Q = reinterpret_cast<affine*>(ctx->small.as<uint8_t>() + 256);
    if (int rc = D::upload(ctx, d_Q, &Q, sizeof(affine))) return rc;

    const affine* curG = d_G;
    const affine* curH = d_H;
    fe fG = Fr::one(), fH = Fr::one();
    bool first = true;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    // stage of the last generator fold (no-fold rounds expand their scalars back to it)
    size_t ns = 0;                        // local length of that stage
    NoFoldParams nf;
    nf.nu = 0;
    const ShardIdx sh{(uint32_t)P, (uint32_t)g};
    // out = L + kappa*R for both generator vectors; GLV (129-step chain) where the curve has the endomorphism
    // Elements with global index in [cross_lo, cross_hi) use the second pair of scalars (kGx, kHx): the partner of such an
    // element lies across the block boundary of a piecewise-geometric factor vector (see below).
    auto fold_uniform = [&](const affine* GL, const affine* GR, affine* Gout, const affine* HL, const affine* HR, affine* Hout, size_t cnt,
                            const fe& kG, const fe& kH, const fe& kGx, const fe& kHx, size_t cross_lo, size_t cross_hi) -> int {
        unsigned grid = (unsigned)((2 * cnt + 127) / 128);
        if constexpr (C::HAS_GLV) {
            GlvSplit sp[4];
            const fe* ks[4] = {&kG, &kH, &kGx, &kHx};
            bool ok = ctx->ipa_glv;
            for (int q = 0; q < 4 && ok; q++) ok = GlvHost<C>::split(*ks[q], sp[q]);
            JsfDigits jd[4];
            bool jsf = ok && ctx->ipa_jsf;
            for (int q = 0; q < 4 && jsf; q++) jsf = jsf_digits(sp[q].k1, sp[q].k2, jd[q]);
            if (jsf) {
                JsfBits b[4];
                for (int q = 0; q < 4; q++) {
                    memcpy(b[q].code, jd[q].code, sizeof(jd[q].code));
                    b[q].neg1 = sp[q].neg1; b[q].neg2 = sp[q].neg2; b[q].top = jd[q].top;
                }
                ipa_fold_points_jsf_kernel<C><<<grid, 128, 0, st>>>(GL, GR, Gout, HL, HR, Hout, cnt, b[0], b[1], b[2], b[3], cross_lo, cross_hi, sh);
                BP_LAUNCH_CHECK(ctx);
                return BP_OK;
            }
            if (ok) {
                GlvBits b[4];
                for (int q = 0; q < 4; q++) {
                    memcpy(b[q].k1, sp[q].k1, 20); memcpy(b[q].k2, sp[q].k2, 20);
                    b[q].neg1 = sp[q].neg1; b[q].neg2 = sp[q].neg2; b[q].top = sp[q].top;
                }
                ipa_fold_points_glv_kernel<C><<<grid, 128, 0, st>>>(GL, GR, Gout, HL, HR, Hout, cnt, b[0], b[1], b[2], b[3], cross_lo, cross_hi, sh);
                BP_LAUNCH_CHECK(ctx);
                return BP_OK;
            }
        }
        ipa_fold_points_uniform_kernel<C><<<grid, 128, 0, st>>>(GL, GR, Gout, HL, HR, Hout, cnt, D::bits(kG), D::bits(kH), D::bits(kGx), D::bits(kHx),
                                                                 cross_lo, cross_hi, sh);
        BP_LAUNCH_CHECK(ctx);
        return BP_OK;
    };
    while (n != 1) {
        auto t_a = now();
        size_t h = n / 2;
        const size_t hl = h / P;          // local half length (0 once the partners live on different ranks)
        const bool nofold = n <= ctx->ipa_nofold_n || hl == 0;
        if (nofold && ns == 0) {
            ns = n / P;                   // generators stay at this stage from now on
            BP_CUDA_TRY(ctx, ctx->ipa_s.reserve((4 * ns + 8) * sizeof(fe)));
            s_all = ctx->ipa_s.as<fe>();
            nf.fG = fG;
            nf.fH = fH;
        }
        const size_t mine = (h + P - 1) / P;      // upper bound of this rank's indices below h
        int blocks = (int)((mine + 127) / 128);
        if (blocks > PREP_BLOCKS) blocks = PREP_BLOCKS;
        MsmJob job;
        if (!nofold) {
            fe *sLG = s_all, *sLH = s_all + hl, *sRG = s_all + 2 * hl, *sRH = s_all + 3 * hl;
            const bool perel = first || geo;
            ipa_prep_kernel<C><<<blocks, 128, 0, st>>>(d_a, d_b, h, sh, perel ? d_Gf : nullptr, perel ? d_Hf : nullptr, fG, fH, sLG, sLH, sRG, sRH, parts);
            BP_LAUNCH_CHECK(ctx);
            // L = <a_L*gR, G_R> + <b_R*hL, H_L> + c_L*Q ; R = <a_R*gL, G_L> + <b_L*hR, H_R> + c_R*Q
            job.add(curG + hl, sLG, hl, 0);
            job.add(curH, sLH, hl, 0);
            job.add(d_Q, d_c, 1, 0);
            job.add(curG, sRG, hl, 1);
            job.add(curH + hl, sRH, hl, 1);
            job.add(d_Q, d_c + 1, 1, 1);
        } else {
            fe *sLG = s_all, *sLH = s_all + ns, *sRG = s_all + 2 * ns, *sRH = s_all + 3 * ns;
            const bool stage0 = first || geo;    // never folded / geometric: the factor vectors still apply per element
            ipa_cross_kernel<C><<<blocks, 128, 0, st>>>(d_a, d_b, h, sh, parts);
            BP_LAUNCH_CHECK(ctx);
            ipa_nofold_scalars_kernel<C><<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(d_a, d_b, ns, n, sh, stage0 ? d_Gf : nullptr, stage0 ? d_Hf : nullptr, nf,
                                                                                      sLG, sLH, sRG, sRH);
            BP_LAUNCH_CHECK(ctx);
            job.add(curG, sLG, ns, 0);
            job.add(curH, sLH, ns, 0);
            job.add(d_Q, d_c, 1, 0);
            job.add(curG, sRG, ns, 1);
            job.add(curH, sRH, ns, 1);
            job.add(d_Q, d_c + 1, 1, 1);
        }
        vec_reduce_partials_kernel<C, 2><<<1, 128, 0, st>>>(parts, blocks, d_c);
        BP_LAUNCH_CHECK(ctx);
        uint8_t out[2][64];
        int ident[2];
        if (int rc = msm_run_job_sharded<C>(ctx, job, 2, out, ident)) return rc;
        auto t_b = now();
        ctx->stage_ms[ST_IPA_MSM] += ms(t_a, t_b);
        affine Lp, Rp;
        memcpy(&Lp, out[0], 64);
        memcpy(&Rp, out[1], 64);
        L_vec.push_back(Lp);
        R_vec.push_back(Rp);
        TP<C>::append_point(t, "L", Lp);
        TP<C>::append_point(t, "R", Rp);
        fe u = TP<C>::challenge_scalar(t, "u");
        fe uinv = Fr::inv(u);                                            // u.inverse().unwrap()
        auto t_c = now();
        ctx->stage_ms[ST_IPA_HOST] += ms(t_b, t_c);
        ipa_fold_scalars_kernel<C><<<(unsigned)((h + 255) / 256), 256, 0, st>>>(d_a, d_b, h, u, uinv);
        BP_LAUNCH_CHECK(ctx);
        if (nofold) {
            if (nf.nu >= 32) return BP_ERR_LEN;
            nf.u[nf.nu] = u;
            nf.uinv[nf.nu] = uinv;
            nf.nu++;
        } else {
            unsigned fgrid = (unsigned)((2 * hl + 127) / 128);
            if (geo) {
                size_t lg_h = 0;
                while (((size_t)1 << lg_h) < h) lg_h++;
                fe rGh = *geo_rG, rHh = *geo_rH;
                for (size_t k = 0; k < lg_h; k++) { rGh = Fr::sqr(rGh); rHh = Fr::sqr(rHh); }
                fe kG = Fr::mul(Fr::sqr(u), rGh), kH = Fr::mul(Fr::sqr(uinv), rHh);
                // block boundary: f[i] jumps by `geo_jump` at index geo_n1, so partners (i, i+h) with i < n1 <= i+h differ by
                // r^h * jump instead of r^h
                fe kGx = kG, kHx = kH;
                size_t cross_lo = 0, cross_hi = 0;
                if (geo_jump && geo_n1 > 0 && geo_n1 < n) {
                    kGx = Fr::mul(kG, *geo_jump);
                    kHx = Fr::mul(kH, *geo_jump);
                    cross_lo = geo_n1 > h ? geo_n1 - h : 0;
                    cross_hi = geo_n1 < h ? geo_n1 : h;
                }
                if (int rc = fold_uniform(curG, curG + hl, wG, curH, curH + hl, wH, hl, kG, kH, kGx, kHx, cross_lo, cross_hi)) return rc;
                fG = Fr::mul(fG, uinv);
                fH = Fr::mul(fH, u);
            } else if (first) {
                // factors folded into the points (inner_product_proof.rs:143-155)
                ipa_fold_points_joint_kernel<C><<<fgrid, 128, 0, st>>>(curG, d_Gf, uinv, u, wG, curH, d_Hf, u, uinv, wH, h, hl, sh);
                BP_LAUNCH_CHECK(ctx);
            } else {
                // u^-1*G_L + u*G_R = u^-1*(G_L + u^2*G_R): the common factor moves into fG (resp. fH)
                fe u2 = Fr::sqr(u), ui2 = Fr::sqr(uinv);
                if (int rc = fold_uniform(curG, curG + hl, wG, curH, curH + hl, wH, hl, u2, ui2, u2, ui2, 0, 0)) return rc;
                fG = Fr::mul(fG, uinv);
                fH = Fr::mul(fH, u);
            }
            if (ctx->timing) { cudaStreamSynchronize(st); ctx->stage_ms[ST_IPA_FOLD] += ms(t_c, now()); }
            curG = wG;
            curH = wH;
            first = false;
        }
        n = h;
    }
    fe ab[2];
    if (int rc = D::download(ctx, &ab[0], d_a, sizeof(fe))) return rc;
    if (int rc = D::download(ctx, &ab[1], d_b, sizeof(fe))) return rc;
    a_out = ab[0];
    b_out = ab[1];
    return BP_OK;
}

// ---- InnerProductProof::verify (inner_product_proof.rs:321-382; test-only in the reference) ----------
// Returns BP_OK iff  a*b*Q + <a*s*Gf, G> + <b*s^-1*Hf, H> - sum u_j^2 L_j - sum u_j^-2 R_j == P.
template <class C>
int ipa_verify(bp_ctx* ctx, Transcript& t, size_t n, const std::vector<affine>& L_vec, const std::vector<affine>& R_vec, const fe& a, const fe& b,
               const fe* d_Gf, const fe* d_Hf, const affine& P, const affine& Q, const affine* d_G, const affine* d_H) {
    using Fr = HostFp<typename C::Fr>;
    using HC = HostCurve<C>;
    using D = Dev<C>;
    size_t lg_n = L_vec.size();
    if (lg_n >= 32 || R_vec.size() != lg_n || n != ((size_t)1 << lg_n)) return BP_ERR_VERIFY;     // :256-264
    t.append_message("dom-sep", (const uint8_t*)"ipp v1", 6);
    t.append_u64("n", n);
    VerifyInputs vin;
    std::vector<fe> tail(1 + 2 * lg_n);
    std::vector<affine> pts(1 + 2 * lg_n);
    fe allinv = Fr::one();
    std::vector<fe> us(lg_n), pre(lg_n);
    fe run = Fr::one();
    for (size_t j = 0; j < lg_n; j++) {
        if (TP<C>::validate_and_append_point(t, "L", L_vec[j])) return BP_ERR_VERIFY;
        if (TP<C>::validate_and_append_point(t, "R", R_vec[j])) return BP_ERR_VERIFY;
        us[j] = TP<C>::challenge_scalar(t, "u");
        pre[j] = run;
        if (!Fr::is_zero(us[j])) run = Fr::mul(run, us[j]);
        pts[1 + j] = L_vec[j];
        pts[1 + lg_n + j] = R_vec[j];
    }
    {   // batch_inversion (:283-288): one inversion, zeros stay zeros
        fe inv = Fr::inv(run);
        allinv = inv;
        for (size_t j = lg_n; j-- > 0;) {
            fe ui = us[j];
            if (!Fr::is_zero(us[j])) { ui = Fr::mul(inv, pre[j]); inv = Fr::mul(inv, us[j]); }
            vin.usq[j] = Fr::sqr(us[j]);
            tail[1 + j] = Fr::neg(vin.usq[j]);                    // neg_u_sq
            tail[1 + lg_n + j] = Fr::neg(Fr::sqr(ui));            // neg_u_inv_sq
        }
    }
    tail[0] = Fr::mul(a, b);
    pts[0] = Q;
    vin.wL = vin.wR = vin.wO = vin.yinvpow = nullptr;
    vin.allinv = allinv; vin.a = a; vin.b = b; vin.x = Fr::zero(); vin.u = Fr::zero(); vin.lg_n = (int)lg_n;
    BP_CUDA_TRY(ctx, ctx->v_g.reserve((n + 1) * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->v_h.reserve((n + 1) * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->v_pts.reserve((pts.size() + 1) * sizeof(affine)));
    BP_CUDA_TRY(ctx, ctx->v_sc.reserve((tail.size() + 1) * sizeof(fe)));
    ipa_verify_scalars_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(vin, n, d_Gf, d_Hf, ctx->v_g.as<fe>(), ctx->v_h.as<fe>());
    BP_LAUNCH_CHECK(ctx);
    D::upload(ctx, ctx->v_sc.p, tail.data(), tail.size() * sizeof(fe));
    if (int rc = D::upload(ctx, ctx->v_pts.p, pts.data(), pts.size() * sizeof(affine))) return rc;
    MsmJob job;
    job.add(d_G, ctx->v_g.as<fe>(), n, 0);
    job.add(d_H, ctx->v_h.as<fe>(), n, 0);
    job.add(ctx->v_pts.as<affine>(), ctx->v_sc.as<fe>(), pts.size(), 0);
    uint8_t o[1][64];
    int id[1] = {0};
    if (int rc = msm_run_job<C>(ctx, job, o, id)) return rc;
    affine got;
    memcpy(&got, o[0], 64);
    const bool p_id = HC::E::is_identity(P);
    if (id[0] || p_id) return (id[0] && p_id) ? BP_OK : BP_ERR_VERIFY;
    return (HC::Fq::eq(got.x, P.x) && HC::Fq::eq(got.y, P.y)) ? BP_OK : BP_ERR_VERIFY;             // :377-381
}

// ---- constraint storage shared by prover and verifier -----------------------------------------------
// Terms are kept in the compact form the device flatten consumes, because the whole store crosses PCIe on every
// prove / verify (4n terms for the chain circuit): a 32-bit sort key (kind << 29 | index) and a 32-bit coefficient
// reference -- 0 = +1, 1 = -1 (the bulk of real circuits), k + 2 = coeff_ex[k] -- instead of 41 bytes per term.
struct ConstraintStore {
    std::vector<uint32_t> key;        // kind << 29 | index
    std::vector<uint32_t> cref;
    std::vector<fe> coeff_ex;
    std::vector<uint32_t> start;      // constraint q covers terms [start[q], start[q+1])
    fe one, minus_one;
    bool overflow = false;            // an index >= 2^29 or more than 2^32 - 3 terms: reported by flatten_device
    // Multi-GPU verifier: rank r only ever needs w_L, w_R, w_O at the generator indices it holds
    // (i = r mod world), so its store keeps only those multiplier terms -- 1/world of the upload, sort and reduce of
    // flatten_device. Committed and constant terms (w_V, w_c) are kept everywhere. The prover keeps everything (its l(x),
    // r(x) vectors are replicated).
    uint32_t shard_rank = 0, shard_world = 1;
    ConstraintStore() { start.push_back(0); }
    void push_term(const Variable& v, const fe& c) {
        if (shard_world > 1 && v.kind >= VAR_MUL_LEFT && v.kind <= VAR_MUL_OUT && (v.idx % shard_world) != shard_rank) return;
        if (v.idx >= (1ull << 29) || key.size() >= 0xFFFFFFF0ull) overflow = true;
        key.push_back(((uint32_t)v.kind << 29) | (uint32_t)(v.idx & 0x1FFFFFFFu));
        if (memcmp(c.v, one.v, 32) == 0) cref.push_back(0u);
        else if (memcmp(c.v, minus_one.v, 32) == 0) cref.push_back(1u);
        else { cref.push_back((uint32_t)coeff_ex.size() + 2u); coeff_ex.push_back(c); }
    }
    void push(const Variable* v, const fe* c, size_t n, const Variable* extra_v = nullptr, const fe* extra_c = nullptr) {
        for (size_t i = 0; i < n; i++) push_term(v[i], c[i]);
        if (extra_v) push_term(*extra_v, *extra_c);
        start.push_back((uint32_t)key.size());
    }
    size_t count() const { return start.size() - 1; }
};

// A Variable that did not come from this constraint system (another prover's, a stale phase-2 one, a forged bp_var)
// panics on an index in the Rust reference; here it would index a_L / wL ... out of bounds. Every term that
// crosses multiply / constrain is checked against the current number of multipliers and commitments.
inline int check_terms(const Variable* v, size_t n, size_t multipliers, size_t commitments) {
    for (size_t i = 0; i < n; i++) {
        switch (v[i].kind) {
            case VAR_MUL_LEFT: case VAR_MUL_RIGHT: case VAR_MUL_OUT:
                if (v[i].idx >= multipliers) return BP_ERR_ARG;
                break;
            case VAR_COMMITTED:
                if (v[i].idx >= commitments) return BP_ERR_ARG;
                break;
            case VAR_ONE: break;
            default: return BP_ERR_ARG;
        }
    }
    return BP_OK;
}

// Device flatten: d_wL/d_wR/d_wO (n each, device) are filled; wV (m) and wc come back to the host.
// `reuse`: batch verification flattens many constraint systems that are usually the same circuit (same terms, same
// coefficients; only z differs). The context remembers the store it uploaded last (an exact host copy, compared with
// memcmp -- no hashing, no false hits) and, on a match, skips the four pageable uploads (4.4 MB at 2^16 multipliers:
// ~0.45 ms, which is what bound bp_batch_verify) and the radix sort, whose result depends on the keys only.
template <class C>
int flatten_device(bp_ctx* ctx, const ConstraintStore& cs, const fe& z, size_t n, size_t m, fe* d_wL, fe* d_wR, fe* d_wO,
                   std::vector<fe>& wV, fe* wc, bool reuse = false) {
    using Fr = HostFp<typename C::Fr>;
    using D = Dev<C>;
    cudaStream_t st = ctx->stream;
    size_t T = cs.key.size(), Q = cs.count(), E = cs.coeff_ex.size();
    if (n >= (1u << 29) || m >= (1u << 29) || cs.overflow) return BP_ERR_LEN;
    wV.assign(m, Fr::zero());
    if (wc) *wc = Fr::zero();
    BP_CUDA_TRY(ctx, cudaMemsetAsync(d_wL, 0, n * sizeof(fe), st));
    BP_CUDA_TRY(ctx, cudaMemsetAsync(d_wR, 0, n * sizeof(fe), st));
    BP_CUDA_TRY(ctx, cudaMemsetAsync(d_wO, 0, n * sizeof(fe), st));
    if (T == 0) return BP_OK;
    BP_CUDA_TRY(ctx, ctx->f_idx.reserve(T * 4));
    BP_CUDA_TRY(ctx, ctx->f_coeff.reserve((E + 1) * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->f_start.reserve((Q + 1) * 4));
    BP_CUDA_TRY(ctx, ctx->f_keys.reserve(T * 4));
    BP_CUDA_TRY(ctx, ctx->f_keys2.reserve(T * 4));
    BP_CUDA_TRY(ctx, ctx->f_perm.reserve(T * 4));
    BP_CUDA_TRY(ctx, ctx->f_perm2.reserve(T * 4));
    BP_CUDA_TRY(ctx, ctx->f_contrib.reserve(T * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->f_sorted.reserve(T * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->f_ukeys.reserve(T * 4));
    BP_CUDA_TRY(ctx, ctx->f_sums.reserve(T * sizeof(fe)));
    BP_CUDA_TRY(ctx, ctx->f_wv.reserve((m + 2) * sizeof(fe) + 16));
    FlattenCache& fc = ctx->flatten_cache;
    const bool hit = reuse && fc.valid && fc.key.size() == T && fc.start.size() == Q + 1 && fc.coeff.size() == E * 8 &&
                     memcmp(fc.key.data(), cs.key.data(), T * 4) == 0 && memcmp(fc.cref.data(), cs.cref.data(), T * 4) == 0 &&
                     memcmp(fc.start.data(), cs.start.data(), (Q + 1) * 4) == 0 &&
                     (E == 0 || memcmp(fc.coeff.data(), cs.coeff_ex.data(), E * sizeof(fe)) == 0);
    fc.valid = false;                                   // the device buffers are about to change (or are confirmed below)
    if (!hit) {
        D::upload(ctx, ctx->f_keys.p, cs.key.data(), T * 4);
        D::upload(ctx, ctx->f_idx.p, cs.cref.data(), T * 4);
        D::upload(ctx, ctx->f_coeff.p, cs.coeff_ex.data(), E * sizeof(fe));
        if (int rc = D::upload(ctx, ctx->f_start.p, cs.start.data(), (Q + 1) * 4)) return rc;
        if (reuse) {
            fc.key = cs.key; fc.cref = cs.cref; fc.start = cs.start;
            fc.coeff.assign(reinterpret_cast<const uint32_t*>(cs.coeff_ex.data()), reinterpret_cast<const uint32_t*>(cs.coeff_ex.data()) + E * 8);
        }
    }
    uint32_t *keys = ctx->f_keys.as<uint32_t>(), *keys2 = ctx->f_keys2.as<uint32_t>(), *perm = ctx->f_perm.as<uint32_t>(), *perm2 = ctx->f_perm2.as<uint32_t>();
    flatten_contrib_kernel<C><<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(ctx->f_idx.as<uint32_t>(), ctx->f_coeff.as<fe>(), ctx->f_start.as<uint32_t>(), Q,
                                                                          D::pow_table(z), perm, ctx->f_contrib.as<fe>());
    BP_LAUNCH_CHECK(ctx);
    size_t tmp1 = 0, tmp2 = 0;
    BP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp1, keys, keys2, perm, perm2, T, 0, 32, st));
    fe* d_wV = ctx->f_wv.as<fe>();
    fe* d_wc = d_wV + m;
    int* d_nruns = reinterpret_cast<int*>(d_wV + m + 1);
    BP_CUDA_TRY(ctx, cub::DeviceReduce::ReduceByKey(nullptr, tmp2, keys2, ctx->f_ukeys.as<uint32_t>(), ctx->f_sorted.as<fe>(), ctx->f_sums.as<fe>(), d_nruns,
                                                    FeAddOp<C>(), (int)T, st));
    BP_CUDA_TRY(ctx, ctx->f_tmp.reserve(tmp1 > tmp2 ? tmp1 : tmp2));
    if (!hit) BP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->f_tmp.p, tmp1, keys, keys2, perm, perm2, T, 0, 32, st));
    flatten_gather_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(ctx->f_contrib.as<fe>(), perm2, T, ctx->f_sorted.as<fe>());
    BP_LAUNCH_CHECK(ctx);
    BP_CUDA_TRY(ctx, cub::DeviceReduce::ReduceByKey(ctx->f_tmp.p, tmp2, keys2, ctx->f_ukeys.as<uint32_t>(), ctx->f_sorted.as<fe>(), ctx->f_sums.as<fe>(), d_nruns,
                                                    FeAddOp<C>(), (int)T, st));
    BP_CUDA_TRY(ctx, cudaMemsetAsync(d_wV, 0, (m + 1) * sizeof(fe), st));
    flatten_scatter_kernel<C><<<(unsigned)((T + 255) / 256), 256, 0, st>>>(ctx->f_ukeys.as<uint32_t>(), ctx->f_sums.as<fe>(), d_nruns, d_wL, d_wR, d_wO, d_wV, d_wc, (uint32_t)n, (uint32_t)m);
    BP_LAUNCH_CHECK(ctx);
    std::vector<fe> back(m + 1);
    if (int rc = D::download(ctx, back.data(), d_wV, (m + 1) * sizeof(fe))) return rc;
    for (size_t i = 0; i < m; i++) wV[i] = back[i];
    if (wc) *wc = back[m];
    fc.valid = reuse;                                   // keys, coefficients, starts, sorted keys and permutation are resident
    return BP_OK;
}

static inline size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }   // 0 -> 1 like Rust

// ---- Prover (src/r1cs/prover.rs) ----------------------------------------------------------------
template <class C>
struct ProverT : ConstraintSystemBase {
    using HC = HostCurve<C>;
    using Fr = HostFp<typename C::Fr>;
    using D = Dev<C>;
    bp_ctx* ctx;
    const GensDev* gens;
    Transcript* transcript;
    std::vector<fe> v, v_blinding, a_L, a_R, a_O;
    ConstraintStore cs;
    std::vector<std::function<int(ConstraintSystemBase&)>> deferred;
    bool has_pending = false, randomizing = false;
    size_t pending = 0;

    ~ProverT() override {                                                                      // prover.rs:74-94 (Drop for Secrets)
        for (auto* w : {&v, &v_blinding, &a_L, &a_R, &a_O})
            if (!w->empty()) explicit_bzero(w->data(), w->size() * sizeof(fe));
    }
    ProverT(bp_ctx* c, const GensDev* g, Transcript* t) : ctx(c), gens(g), transcript(t) {     // prover.rs:291-308
        t->append_message("dom-sep", (const uint8_t*)"r1cs v1", 7);
        cs.one = Fr::one();
        cs.minus_one = Fr::neg(Fr::one());
    }
    fe eval(const Variable* vars, const fe* coeffs, size_t n) const {                          // prover.rs:399-414
        fe tot = Fr::zero();
        for (size_t k = 0; k < n; k++) {
            fe val;
            switch (vars[k].kind) {
                case VAR_MUL_LEFT: val = a_L[vars[k].idx]; break;
                case VAR_MUL_RIGHT: val = a_R[vars[k].idx]; break;
                case VAR_MUL_OUT: val = a_O[vars[k].idx]; break;
                case VAR_COMMITTED: val = v[vars[k].idx]; break;
                case VAR_ONE: val = Fr::one(); break;
                default: val = Fr::zero();
            }
            tot = Fr::add(tot, Fr::mul(coeffs[k], val));
        }
        return tot;
    }
    int multiply(const Variable* lv, const fe* lc, size_t ln, const Variable* rv, const fe* rc, size_t rn, Variable out[3]) override {   // :103-133
        if (int rc_ = check_terms(lv, ln, a_L.size(), v.size())) return rc_;
        if (int rc_ = check_terms(rv, rn, a_L.size(), v.size())) return rc_;
        fe l = eval(lv, lc, ln), r = eval(rv, rc, rn);
        fe o = Fr::mul(l, r);
        uint64_t i = a_L.size();
        out[0] = {VAR_MUL_LEFT, i}; out[1] = {VAR_MUL_RIGHT, i}; out[2] = {VAR_MUL_OUT, i};
        a_L.push_back(l); a_R.push_back(r); a_O.push_back(o);
        fe minus_one = Fr::neg(Fr::one());
        cs.push(lv, lc, ln, &out[0], &minus_one);
        cs.push(rv, rc, rn, &out[1], &minus_one);
        return BP_OK;
    }
    int allocate(const fe* assignment, Variable* out) override {                               // :135-157
        if (!assignment) return BP_ERR_MISSING;
        if (!has_pending) {
            uint64_t i = a_L.size();
            has_pending = true; pending = i;
            a_L.push_back(*assignment); a_R.push_back(Fr::zero()); a_O.push_back(Fr::zero());
            *out = {VAR_MUL_LEFT, i};
        } else {
            has_pending = false;
            a_R[pending] = *assignment;
            a_O[pending] = Fr::mul(a_L[pending], a_R[pending]);
            *out = {VAR_MUL_RIGHT, pending};
        }
        return BP_OK;
    }
    int allocate_multiplier(const fe* l, const fe* r, Variable out[3]) override {               // :159-183
        if (!l || !r) return BP_ERR_MISSING;
        uint64_t i = a_L.size();
        out[0] = {VAR_MUL_LEFT, i}; out[1] = {VAR_MUL_RIGHT, i}; out[2] = {VAR_MUL_OUT, i};
        a_L.push_back(*l); a_R.push_back(*r); a_O.push_back(Fr::mul(*l, *r));
        return BP_OK;
    }
    int constrain(const Variable* vv, const fe* c, size_t n) override {                         // :189-193
        if (int rc_ = check_terms(vv, n, a_L.size(), v.size())) return rc_;
        cs.push(vv, c, n);
        return BP_OK;
    }
    size_t multipliers_len() const override { return a_L.size(); }
    int challenge_scalar(const char* label, fe* out) override {                                 // :262-267
        if (!randomizing) return BP_ERR_ARG;
        *out = TP<C>::challenge_scalar(*transcript, label);
        return BP_OK;
    }
    int specify_randomized_constraints(std::function<int(ConstraintSystemBase&)> cb) override { deferred.push_back(std::move(cb)); return BP_OK; }
    // commit (prover.rs:327-341): V = v*B + v_blinding*B_blinding
    int commit(const fe& val, const fe& blind, affine& V, Variable& var) {
        uint64_t i = v.size();
        v.push_back(val);
        v_blinding.push_back(blind);
        V = HC::add(HC::mul(gens->B, val), HC::mul(gens->B_blinding, blind));                   // generators.rs:39-44
        TP<C>::append_point(*transcript, "V", V);
        var = {VAR_COMMITTED, i};
        return BP_OK;
    }
    // m commitments at once: the scalar multiplications run on the GPU (pedersen_commit_kernel), the
    // transcript appends stay in commit order on the host -- same V_i, same transcript as m commit() calls
    int commit_batch(const fe* vals, const fe* blinds, size_t m, affine* V_out, Variable* vars) {
        if (m == 0) return BP_OK;
        // context-owned grow-only buffers: a cudaMalloc/cudaFree pair per call synchronises the device and was measured at
        // up to ~100 ms in a process that holds many GB of other allocations (bench.py after the 2^24 MSM)
        DevBuf &dv = ctx->c_v, &db = ctx->c_b, &dout = ctx->c_out;
        BP_CUDA_TRY(ctx, dv.reserve(m * sizeof(fe)));
        BP_CUDA_TRY(ctx, db.reserve(m * sizeof(fe)));
        BP_CUDA_TRY(ctx, dout.reserve(m * sizeof(affine)));
        D::upload(ctx, dv.p, vals, m * sizeof(fe));
        if (int rc = D::upload(ctx, db.p, blinds, m * sizeof(fe))) return rc;
        if (ctx->pedersen_table) {
            if (!gens->has_pc_table) {
                DevBuf win;
                BP_CUDA_TRY(ctx, win.reserve(64 * sizeof(affine)));
                cudaError_t e = gens->pc_table.reserve((size_t)2 * 32 * 256 * sizeof(affine));
                if (e != cudaSuccess) { win.release(); ctx->err = "pc_table"; return BP_ERR_CUDA; }
                pedersen_window_kernel<C><<<1, 64, 0, ctx->stream>>>(gens->B, gens->B_blinding, win.as<affine>());
                ctx->launches++;
                pedersen_table_kernel<C><<<(2 * 32 * 256) / 128, 128, 0, ctx->stream>>>(win.as<affine>(), gens->pc_table.template as<affine>());
                ctx->launches++;
                e = cudaStreamSynchronize(ctx->stream);
                win.release();
                if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return BP_ERR_CUDA; }
                gens->has_pc_table = true;
            }
            pedersen_commit_table_kernel<C><<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(gens->pc_table.template as<affine>(), dv.as<fe>(),
                                                                                                 db.as<fe>(), dout.as<affine>(), m);
        } else {
            affine BBb = HC::add(gens->B, gens->B_blinding);
            pedersen_commit_kernel<C><<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(gens->B, gens->B_blinding, BBb, dv.as<fe>(), db.as<fe>(),
                                                                                           dout.as<affine>(), m);
        }
        BP_LAUNCH_CHECK(ctx);
        int rc_dl = D::download(ctx, V_out, dout.p, m * sizeof(affine));
        cudaMemsetAsync(dv.p, 0, m * sizeof(fe), ctx->stream);        // values and blindings do not outlive the call on the device
        cudaMemsetAsync(db.p, 0, m * sizeof(fe), ctx->stream);
        if (rc_dl) return rc_dl;
        for (size_t i = 0; i < m; i++) {
            vars[i] = {VAR_COMMITTED, (uint64_t)v.size()};
            v.push_back(vals[i]);
            v_blinding.push_back(blinds[i]);
            TP<C>::append_point(*transcript, "V", V_out[i]);
        }
        return BP_OK;
    }
    int create_randomized_constraints() {                                                       // :418-441
        has_pending = false;
        if (deferred.empty()) {
            transcript->append_message("dom-sep", (const uint8_t*)"r1cs-1phase", 11);
        } else {
            transcript->append_message("dom-sep", (const uint8_t*)"r1cs-2phase", 11);
            auto cbs = std::move(deferred);
            deferred.clear();
            randomizing = true;
            for (auto& cb : cbs)
                if (int rc = cb(*this)) { randomizing = false; return rc; }
            randomizing = false;
        }
        return BP_OK;
    }

    // three vector commitments over generators [off, off+cnt) in one batched MSM (prover.rs:516-559 / 604-649)
    // `which`: bit 0 = A_I and A_O (out[0], out[1]), bit 1 = S (out[2]); 3 = all three in one batch. Large one-phase
    // circuits commit A_I/A_O while the host is still drawing s_L, s_R (see prove()).
    int commit_phase(size_t off, size_t cnt, const fe bl[3], const fe* d_aL, const fe* d_aR, const fe* d_aO, const fe* d_sL, const fe* d_sR,
                     affine out[3], int which = 3) {
        fe* d_bl = ctx->small.as<fe>() + 16;
        if (int rc = D::upload(ctx, d_bl, bl, 3 * sizeof(fe))) return rc;
        const affine* Bb = gens->pc.template as<affine>() + 1;
        // sharded contexts: this rank's generators of [off, off+cnt) against the replicated scalar vectors; the
        // blinding terms are added once (rank 0); the partial points are all-gathered and summed
        const bool lead = ctx->rank == 0;
        MsmJob job;
        int nm = 0, slot[3] = {-1, -1, -1};
        if (which & 1) {
            slot[0] = nm++; slot[1] = nm++;
            if (lead) job.add(Bb, d_bl, 1, slot[0]);                                                   // A_I
            gens->add_range(job, gens->G, d_aL, off, cnt, slot[0]); gens->add_range(job, gens->H, d_aR, off, cnt, slot[0]);
            if (lead) job.add(Bb, d_bl + 1, 1, slot[1]);                                               // A_O
            gens->add_range(job, gens->G, d_aO, off, cnt, slot[1]);
        }
        if (which & 2) {
            slot[2] = nm++;
            if (lead) job.add(Bb, d_bl + 2, 1, slot[2]);                                               // S
            gens->add_range(job, gens->G, d_sL, off, cnt, slot[2]); gens->add_range(job, gens->H, d_sR, off, cnt, slot[2]);
        }
        uint8_t o[3][64];
        int id[3];
        if (int rc = msm_run_job_sharded<C>(ctx, job, nm, o, id)) return rc;
        for (int k = 0; k < 3; k++)
            if (slot[k] >= 0) memcpy(&out[k], o[slot[k]], 64);
        return BP_OK;
    }

    // prove_and_return_transcript (prover.rs:454-831)
    int prove(Rng& prng, ProofT<C>& proof) {
        Transcript& t = *transcript;
        cudaStream_t st = ctx->stream;
        StageTimer tm(ctx);
        t.append_u64("m", v.size());                                                            // :466
        std::vector<std::vector<uint8_t>> wit;
        for (auto& vb : v_blinding) { std::vector<uint8_t> b(32); HC::scalar_to_bytes(vb, b.data()); wit.push_back(b); }
        TranscriptRng rng = t.make_rng("v_blinding", wit, prng);                                // :483-494
        size_t n1 = a_L.size();
        if (gens->capacity < n1) return BP_ERR_GENS;                                            // :499-501
        fe bl1[3], bl2[3] = {Fr::zero(), Fr::zero(), Fr::zero()}, tb[6];
        for (int k = 0; k < 3; k++) bl1[k] = HC::scalar_rand(rng);                              // :506-508
        std::vector<fe> s_L(n1), s_R(n1);
        // Secrets (mirrors `impl Drop for Secrets` and the clearing at prover.rs:74-94,805-812): the blinding scalars and
        // vectors on the host and every witness-derived device buffer are zeroed when prove() leaves, on the error
        // paths too. Declared before the draw thread's joiner, so it runs after the thread has stopped writing.
        struct Wiper {
            bp_ctx* ctx; fe *bl1, *bl2, *tb; std::vector<fe>*sl, *sr;
            ~Wiper() {
                DevBuf* sec[] = {&ctx->p_aL, &ctx->p_aR, &ctx->p_aO, &ctx->p_sL, &ctx->p_sR, &ctx->p_l, &ctx->p_r, &ctx->ipa_s};
                for (auto* b : sec) if (b->p) cudaMemsetAsync(b->p, 0, b->cap, ctx->stream);
                explicit_bzero(bl1, 3 * sizeof(fe)); explicit_bzero(bl2, 3 * sizeof(fe)); explicit_bzero(tb, 6 * sizeof(fe));
                if (!sl->empty()) explicit_bzero(sl->data(), sl->size() * sizeof(fe));   // one bulk clear the compiler may not elide
                if (!sr->empty()) explicit_bzero(sr->data(), sr->size() * sizeof(fe));
            }
        } wiper{ctx, bl1, bl2, tb, &s_L, &s_R};
        // The 8*n1 dependent Keccak permutations behind s_L, s_R are the longest stage of a large proof and need only
        // the host: for large one-phase parts they run on a second host thread while this one uploads a_L, a_R, a_O and
        // commits A_I, A_O (which do not depend on them); S follows when the draws are done. Same draws, same order.
        const bool overlap = n1 >= 2048;
        std::thread drawer;
        auto draw = [&] {
            HC::scalar_rand_bulk(rng, s_L.data(), n1);                                          // :510-513
            HC::scalar_rand_bulk(rng, s_R.data(), n1);
        };
        if (overlap) drawer = std::thread(draw);
        else draw();
        struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{drawer};   // also on the error paths
        if (!overlap) tm.lap(ST_RNG);
        BP_CUDA_TRY(ctx, ctx->small.reserve(4096));
        // device vectors sized for phase 1; grown after the randomised phase
        auto need = [&](size_t n_) -> int {
            DevBuf* bufs[] = {&ctx->p_aL, &ctx->p_aR, &ctx->p_aO, &ctx->p_sL, &ctx->p_sR};
            for (auto* b : bufs) BP_CUDA_TRY(ctx, b->reserve((n_ + 1) * sizeof(fe)));
            return BP_OK;
        };
        if (int rc = need(n1)) return rc;
        fe *d_aL = ctx->p_aL.as<fe>(), *d_aR = ctx->p_aR.as<fe>(), *d_aO = ctx->p_aO.as<fe>(), *d_sL = ctx->p_sL.as<fe>(), *d_sR = ctx->p_sR.as<fe>();
        D::upload(ctx, d_aL, a_L.data(), n1 * sizeof(fe)); D::upload(ctx, d_aR, a_R.data(), n1 * sizeof(fe));
        if (int rc = D::upload(ctx, d_aO, a_O.data(), n1 * sizeof(fe))) return rc;
        affine c1[3];
        if (overlap) {
            if (int rc = commit_phase(0, n1, bl1, d_aL, d_aR, d_aO, d_sL, d_sR, c1, 1)) return rc;   // A_I, A_O under the draws
            drawer.join();
            tm.lap(ST_RNG);
        }
        D::upload(ctx, d_sL, s_L.data(), n1 * sizeof(fe));
        if (int rc = D::upload(ctx, d_sR, s_R.data(), n1 * sizeof(fe))) return rc;
        tm.lap(ST_UPLOAD);
        if (int rc = commit_phase(0, n1, bl1, d_aL, d_aR, d_aO, d_sL, d_sR, c1, overlap ? 2 : 3)) return rc;
        tm.lap(ST_COMMIT);
        proof.A_I1 = c1[0]; proof.A_O1 = c1[1]; proof.S1 = c1[2];
        TP<C>::append_point(t, "A_I1", proof.A_I1);                                             // :561-564
        TP<C>::append_point(t, "A_O1", proof.A_O1);
        TP<C>::append_point(t, "S1", proof.S1);
        if (int rc = create_randomized_constraints()) return rc;                                // :567
        size_t n = a_L.size(), n2 = n - n1, padded_n = next_pow2(n), pad = padded_n - n;
        if (gens->capacity < padded_n) return BP_ERR_GENS;                                      // :577-579
        if (n2 > 0) for (int k = 0; k < 3; k++) bl2[k] = HC::scalar_rand(rng);                   // :585-597
        s_L.resize(n); s_R.resize(n);
        HC::scalar_rand_bulk(rng, s_L.data() + n1, n - n1);                                     // :599-602
        HC::scalar_rand_bulk(rng, s_R.data() + n1, n - n1);
        tm.lap(ST_RNG);
        if (n2 > 0) {
            // grow (contents of phase 1 are re-uploaded: the arena may move)
            if (int rc = need(n)) return rc;
            d_aL = ctx->p_aL.as<fe>(); d_aR = ctx->p_aR.as<fe>(); d_aO = ctx->p_aO.as<fe>(); d_sL = ctx->p_sL.as<fe>(); d_sR = ctx->p_sR.as<fe>();
            D::upload(ctx, d_aL, a_L.data(), n * sizeof(fe)); D::upload(ctx, d_aR, a_R.data(), n * sizeof(fe));
            D::upload(ctx, d_aO, a_O.data(), n * sizeof(fe)); D::upload(ctx, d_sL, s_L.data(), n * sizeof(fe));
            if (int rc = D::upload(ctx, d_sR, s_R.data(), n * sizeof(fe))) return rc;
            affine c2[3];
            if (int rc = commit_phase(n1, n2, bl2, d_aL + n1, d_aR + n1, d_aO + n1, d_sL + n1, d_sR + n1, c2)) return rc;
            proof.A_I2 = c2[0]; proof.A_O2 = c2[1]; proof.S2 = c2[2];
            tm.lap(ST_COMMIT);
        } else {
            proof.A_I2 = proof.A_O2 = proof.S2 = HC::E::affine_identity();                      // :651-655
        }
        TP<C>::append_point(t, "A_I2", proof.A_I2);                                             // :658-661
        TP<C>::append_point(t, "A_O2", proof.A_O2);
        TP<C>::append_point(t, "S2", proof.S2);
        fe y = TP<C>::challenge_scalar(t, "y"), z = TP<C>::challenge_scalar(t, "z");            // :665-667
        std::vector<fe> wV;
        DevBuf* wb0[] = {&ctx->p_wL, &ctx->p_wR, &ctx->p_wO};
        for (auto* b : wb0) BP_CUDA_TRY(ctx, b->reserve((n + 1) * sizeof(fe)));
        if (int rc = flatten_device<C>(ctx, cs, z, n, v.size(), ctx->p_wL.as<fe>(), ctx->p_wR.as<fe>(), ctx->p_wO.as<fe>(), wV, nullptr)) return rc;   // :669
        tm.lap(ST_FLATTEN);
        fe y_inv = Fr::inv(y);                                                                  // :675
        // device: w vectors, power tables, l/r/t
        DevBuf* wb[] = {&ctx->p_wL, &ctx->p_wR, &ctx->p_wO};
        for (auto* b : wb) BP_CUDA_TRY(ctx, b->reserve((n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_ypow.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_yinv.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_l.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_r.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_Gf.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_Hf.reserve((padded_n + 1) * sizeof(fe)));
        const int TP_BLOCKS = 296;
        BP_CUDA_TRY(ctx, ctx->ipa_parts.reserve((size_t)(6 * TP_BLOCKS + 8) * sizeof(fe)));
        if (int rc = D::pow_vec(ctx, y, ctx->p_ypow.as<fe>(), padded_n)) return rc;              // exp_iter(y), util.rs:35-58
        if (int rc = D::pow_vec(ctx, y_inv, ctx->p_yinv.as<fe>(), padded_n)) return rc;          // :676-678
        LrInputs in{d_aL, d_aR, d_aO, d_sL, d_sR, ctx->p_wL.as<fe>(), ctx->p_wR.as<fe>(), ctx->p_wO.as<fe>(), ctx->p_ypow.as<fe>(), ctx->p_yinv.as<fe>()};
        fe tc[6];
        for (auto& x : tc) x = Fr::zero();
        if (n > 0) {
            int blocks = (int)((n + 127) / 128);
            if (blocks > TP_BLOCKS) blocks = TP_BLOCKS;
            r1cs_tpoly_kernel<C><<<blocks, 128, 0, st>>>(in, n, ctx->ipa_parts.as<fe>());       // :684-703
            BP_LAUNCH_CHECK(ctx);
            fe* d_t = ctx->small.as<fe>() + 32;
            vec_reduce_partials_kernel<C, 6><<<1, 128, 0, st>>>(ctx->ipa_parts.as<fe>(), blocks, d_t);
            BP_LAUNCH_CHECK(ctx);
            if (int rc = D::download(ctx, tc, d_t, 6 * sizeof(fe))) return rc;
        }
        tm.lap(ST_VEC);
        tb[0] = HC::scalar_rand(rng);                                                           // t_1_blinding :705
        for (int k = 2; k < 6; k++) tb[k] = HC::scalar_rand(rng);                               // t_3..t_6   :706-709
        // T_i = t_i*B + tb_i*B_blinding for i in {1,3,4,5,6}: five 2-term MSMs in one batch (:711-715)
        {
            fe sc[10];
            const int ids[5] = {0, 2, 3, 4, 5};
            for (int k = 0; k < 5; k++) { sc[2 * k] = tc[ids[k]]; sc[2 * k + 1] = tb[ids[k]]; }
            fe* d_sc = ctx->small.as<fe>() + 48;
            if (int rc = D::upload(ctx, d_sc, sc, sizeof(sc))) return rc;
            MsmJob job;
            for (int k = 0; k < 5; k++) job.add(gens->pc.template as<affine>(), d_sc + 2 * k, 2, k);
            uint8_t o[5][64];
            int id[5];
            if (int rc = msm_run_job<C>(ctx, job, o, id)) return rc;
            affine* Ts[5] = {&proof.T_1, &proof.T_3, &proof.T_4, &proof.T_5, &proof.T_6};
            for (int k = 0; k < 5; k++) memcpy(Ts[k], o[k], 64);
        }
        tm.lap(ST_TCOMMIT);
        TP<C>::append_point(t, "T_1", proof.T_1); TP<C>::append_point(t, "T_3", proof.T_3);     // :717-722
        TP<C>::append_point(t, "T_4", proof.T_4); TP<C>::append_point(t, "T_5", proof.T_5);
        TP<C>::append_point(t, "T_6", proof.T_6);
        fe u = TP<C>::challenge_scalar(t, "u"), x = TP<C>::challenge_scalar(t, "x");            // :724-725
        tb[1] = Fr::zero();                                                                     // t_2_blinding :729-733
        for (size_t i = 0; i < wV.size(); i++) tb[1] = Fr::add(tb[1], Fr::mul(v_blinding[i], wV[i]));
        auto poly6 = [&](const fe* c) {                                                         // util.rs:107-109
            fe acc = c[5];
            for (int k = 4; k >= 0; k--) acc = Fr::add(c[k], Fr::mul(x, acc));
            return Fr::mul(x, acc);
        };
        proof.t_x = poly6(tc);                                                                  // :744-745
        proof.t_x_blinding = poly6(tb);
        r1cs_lr_eval_kernel<C><<<(unsigned)((padded_n + 255) / 256), 256, 0, st>>>(in, n, padded_n, n1, x, u, ctx->p_l.as<fe>(), ctx->p_r.as<fe>(),
                                                                                  ctx->p_Gf.as<fe>(), ctx->p_Hf.as<fe>());   // :746-756, :781-789
        BP_LAUNCH_CHECK(ctx);
        (void)pad;
        fe i_bl = Fr::add(bl1[0], Fr::mul(u, bl2[0])), o_bl = Fr::add(bl1[1], Fr::mul(u, bl2[1])), s_bl = Fr::add(bl1[2], Fr::mul(u, bl2[2]));
        proof.e_blinding = Fr::mul(x, Fr::add(i_bl, Fr::mul(x, Fr::add(o_bl, Fr::mul(x, s_bl)))));   // :758-762
        TP<C>::append_scalar(t, "t_x", proof.t_x);                                              // :764-774
        TP<C>::append_scalar(t, "t_x_blinding", proof.t_x_blinding);
        TP<C>::append_scalar(t, "e_blinding", proof.e_blinding);
        fe w = TP<C>::challenge_scalar(t, "w");                                                 // :777-779
        affine Q = HC::mul(gens->B, w);
        tm.lap(ST_VEC);
        // G_factors = [1]*n1 ++ [u]*(n2+pad), H_factors[i] = y^-i * G_factors[i] (:781-789): geometric with one jump (by u, at n1)
        const bool geo = ctx->ipa_geo;
        const fe geo_rG = Fr::one();
        int rc = ipa_create<C>(ctx, t, Q, ctx->p_Gf.as<fe>(), ctx->p_Hf.as<fe>(), gens->G.template as<affine>(), gens->H.template as<affine>(),
                               ctx->p_l.as<fe>(), ctx->p_r.as<fe>(), padded_n, proof.L_vec, proof.R_vec, proof.a, proof.b,
                               geo ? &geo_rG : nullptr, geo ? &y_inv : nullptr, n1, &u);                                // :791-800
        tm.lap(ST_IPA);
        tm.lap(ST_TAIL);          // the wiper's clears are queued on the stream as prove() returns
        return rc;
    }
};

// ---- Verifier (src/r1cs/verifier.rs) ------------------------------------------------------------
template <class C>
struct VerifierT : ConstraintSystemBase {
    using HC = HostCurve<C>;
    using Fr = HostFp<typename C::Fr>;
    using D = Dev<C>;
    bp_ctx* ctx;
    Transcript* transcript;
    size_t num_vars = 0;
    std::vector<affine> V;
    ConstraintStore cs;
    std::vector<std::function<int(ConstraintSystemBase&)>> deferred;
    bool has_pending = false, randomizing = false;
    size_t pending = 0;

    VerifierT(bp_ctx* c, Transcript* t) : ctx(c), transcript(t) {                               // :252-263
        t->append_message("dom-sep", (const uint8_t*)"r1cs v1", 7);
        cs.one = Fr::one();
        cs.minus_one = Fr::neg(Fr::one());
        cs.shard_rank = (uint32_t)c->rank;
        cs.shard_world = (uint32_t)c->world;
    }
    int multiply(const Variable* lv, const fe* lc, size_t ln, const Variable* rv, const fe* rc, size_t rn, Variable out[3]) override {   // :74-98
        if (int rc_ = check_terms(lv, ln, num_vars, V.size())) return rc_;
        if (int rc_ = check_terms(rv, rn, num_vars, V.size())) return rc_;
        uint64_t i = num_vars++;
        out[0] = {VAR_MUL_LEFT, i}; out[1] = {VAR_MUL_RIGHT, i}; out[2] = {VAR_MUL_OUT, i};
        fe minus_one = Fr::neg(Fr::one());
        cs.push(lv, lc, ln, &out[0], &minus_one);
        cs.push(rv, rc, rn, &out[1], &minus_one);
        return BP_OK;
    }
    int allocate(const fe*, Variable* out) override {                                           // :100-116
        if (!has_pending) { uint64_t i = num_vars++; has_pending = true; pending = i; *out = {VAR_MUL_LEFT, i}; }
        else { has_pending = false; *out = {VAR_MUL_RIGHT, pending}; }
        return BP_OK;
    }
    int allocate_multiplier(const fe*, const fe*, Variable out[3]) override {                   // :118-137
        uint64_t i = num_vars++;
        out[0] = {VAR_MUL_LEFT, i}; out[1] = {VAR_MUL_RIGHT, i}; out[2] = {VAR_MUL_OUT, i};
        return BP_OK;
    }
    int constrain(const Variable* vv, const fe* c, size_t n) override {
        if (int rc_ = check_terms(vv, n, num_vars, V.size())) return rc_;
        cs.push(vv, c, n);
        return BP_OK;
    }
    size_t multipliers_len() const override { return num_vars; }
    int challenge_scalar(const char* label, fe* out) override {
        if (!randomizing) return BP_ERR_ARG;
        *out = TP<C>::challenge_scalar(*transcript, label);
        return BP_OK;
    }
    int specify_randomized_constraints(std::function<int(ConstraintSystemBase&)> cb) override { deferred.push_back(std::move(cb)); return BP_OK; }
    int commit(const affine& commitment, Variable& var) {                                       // :279-287
        uint64_t i = V.size();
        V.push_back(commitment);
        TP<C>::append_point(*transcript, "V", commitment);
        var = {VAR_COMMITTED, i};
        return BP_OK;
    }
    int create_randomized_constraints() {                                                       // :353-376
        has_pending = false;
        if (deferred.empty()) {
            transcript->append_message("dom-sep", (const uint8_t*)"r1cs-1phase", 11);
        } else {
            transcript->append_message("dom-sep", (const uint8_t*)"r1cs-2phase", 11);
            auto cbs = std::move(deferred);
            deferred.clear();
            randomizing = true;
            for (auto& cb : cbs)
                if (int rc = cb(*this)) { randomizing = false; return rc; }
            randomizing = false;
        }
        return BP_OK;
    }

    // Result of verification_scalars (verifier.rs:394-541): g/h scalars stay on the device
    // (d_g, d_h, padded_n each); head = [B, B_blinding] scalars; tail = scalars of
    // A_I1..S2, V, T_1..T_6, L, R in the reference's order.
    struct Scalars {
        size_t padded_n = 0;
        fe head[2];
        std::vector<fe> tail;
        fe *g = nullptr, *h = nullptr;   // context-owned device buffers (valid until the next verification_scalars)
    };

    // verification_scalars (verifier.rs:394-541) in three steps, so that a batch can derive the inner-product challenges
    // of all its proofs in one launch (transcript_dev.cuh) between the first and the last:
    //   vs_head        the transcript through the challenge `w` (host; runs the randomised-phase callbacks)
    //   vs_challenges  the IPA challenges u_j, their inverses and `r` (host here; device for batches)
    //   vs_tail        flatten, the g / h scalar kernel, the tail scalars
    struct Head {
        fe y, z, u, x, w, r;
        size_t n1 = 0, n = 0, padded_n = 0, lg_n = 0;
        std::vector<fe> ch, ch_inv;      // u_j and u_j^-1 (zeros stay zeros)
    };
    int vs_head(const ProofT<C>& proof, const GensDev& gens, Head& h) {
        Transcript& t = *transcript;
        t.append_u64("m", V.size());                                                            // :404
        h.n1 = num_vars;
        if (TP<C>::validate_and_append_point(t, "A_I1", proof.A_I1)) return BP_ERR_VERIFY;      // :407-409
        if (TP<C>::validate_and_append_point(t, "A_O1", proof.A_O1)) return BP_ERR_VERIFY;
        if (TP<C>::validate_and_append_point(t, "S1", proof.S1)) return BP_ERR_VERIFY;
        if (int rc = create_randomized_constraints()) return rc;                                // :412
        h.n = num_vars;
        h.padded_n = next_pow2(h.n);
        if (gens.capacity < h.padded_n) return BP_ERR_GENS;                                     // :425-427
        TP<C>::append_point(t, "A_I2", proof.A_I2);                                             // :430-432
        TP<C>::append_point(t, "A_O2", proof.A_O2);
        TP<C>::append_point(t, "S2", proof.S2);
        h.y = TP<C>::challenge_scalar(t, "y");                                                  // :434-436
        h.z = TP<C>::challenge_scalar(t, "z");
        if (TP<C>::validate_and_append_point(t, "T_1", proof.T_1)) return BP_ERR_VERIFY;        // :438-442
        if (TP<C>::validate_and_append_point(t, "T_3", proof.T_3)) return BP_ERR_VERIFY;
        if (TP<C>::validate_and_append_point(t, "T_4", proof.T_4)) return BP_ERR_VERIFY;
        if (TP<C>::validate_and_append_point(t, "T_5", proof.T_5)) return BP_ERR_VERIFY;
        if (TP<C>::validate_and_append_point(t, "T_6", proof.T_6)) return BP_ERR_VERIFY;
        h.u = TP<C>::challenge_scalar(t, "u");                                                  // :444-445
        h.x = TP<C>::challenge_scalar(t, "x");
        TP<C>::append_scalar(t, "t_x", proof.t_x);                                              // :447-457
        TP<C>::append_scalar(t, "t_x_blinding", proof.t_x_blinding);
        TP<C>::append_scalar(t, "e_blinding", proof.e_blinding);
        h.w = TP<C>::challenge_scalar(t, "w");                                                  // :459
        // InnerProductProof::verification_scalars (inner_product_proof.rs:244-314): shape checks
        h.lg_n = proof.L_vec.size();
        if (h.lg_n >= 32 || proof.R_vec.size() != h.lg_n || h.padded_n != ((size_t)1 << h.lg_n)) return BP_ERR_VERIFY;   // :256-264
        return BP_OK;
    }
    int vs_challenges_host(const ProofT<C>& proof, Head& h) {
        Transcript& t = *transcript;
        const size_t lg_n = h.lg_n;
        t.append_message("dom-sep", (const uint8_t*)"ipp v1", 6);
        t.append_u64("n", h.padded_n);
        h.ch.assign(lg_n, Fr::zero());
        h.ch_inv.assign(lg_n, Fr::zero());
        for (size_t j = 0; j < lg_n; j++) {                                                     // :271-277
            if (TP<C>::validate_and_append_point(t, "L", proof.L_vec[j])) return BP_ERR_VERIFY;
            if (TP<C>::validate_and_append_point(t, "R", proof.R_vec[j])) return BP_ERR_VERIFY;
            h.ch[j] = TP<C>::challenge_scalar(t, "u");
        }
        // ark_ff::batch_inversion (:283-288): Montgomery's trick, zeros stay zeros -- one field inversion for all rounds
        // (sixteen separate Fermat inversions were 0.2 ms of a 2.1 ms verification at 2^16 multipliers)
        std::vector<fe> pre(lg_n);
        fe run = Fr::one();
        for (size_t j = 0; j < lg_n; j++) {
            pre[j] = run;
            if (!Fr::is_zero(h.ch[j])) run = Fr::mul(run, h.ch[j]);
        }
        fe inv = Fr::inv(run);
        for (size_t j = lg_n; j-- > 0;) {
            if (Fr::is_zero(h.ch[j])) { h.ch_inv[j] = h.ch[j]; continue; }
            h.ch_inv[j] = Fr::mul(inv, pre[j]);
            inv = Fr::mul(inv, h.ch[j]);
        }
        // r: challenge on a CLONE of the transcript (verifier.rs:516-519)
        Transcript tc = t;
        h.r = TP<C>::challenge_scalar(tc, "r");
        return BP_OK;
    }
    int verification_scalars(const ProofT<C>& proof, const GensDev& gens, Scalars& out, bool reuse_store = false) {
        Head h;
        if (int rc = vs_head(proof, gens, h)) return rc;
        if (int rc = vs_challenges_host(proof, h)) return rc;
        return vs_tail(proof, gens, h, out, reuse_store);
    }
    int vs_tail(const ProofT<C>& proof, const GensDev& gens, Head& hd, Scalars& out, bool reuse_store = false) {
        cudaStream_t st = ctx->stream;
        const fe &y = hd.y, &z = hd.z, &u = hd.u, &x = hd.x, &w = hd.w, &r = hd.r;
        const size_t n1 = hd.n1, n = hd.n, padded_n = hd.padded_n, lg_n = hd.lg_n;
        (void)gens;
        std::vector<fe> wV;
        fe wc;
        DevBuf* wb0[] = {&ctx->p_wL, &ctx->p_wR, &ctx->p_wO};
        for (auto* bf : wb0) BP_CUDA_TRY(ctx, bf->reserve((n + 1) * sizeof(fe)));
        if (int rc = flatten_device<C>(ctx, cs, z, n, V.size(), ctx->p_wL.as<fe>(), ctx->p_wR.as<fe>(), ctx->p_wO.as<fe>(), wV, &wc, reuse_store)) return rc;   // :462
        std::vector<fe> ch = hd.ch, ch_inv = hd.ch_inv;
        fe allinv = Fr::one();                                                                  // product of the inverses of the non-zero challenges
        for (size_t j = 0; j < lg_n; j++)
            if (!Fr::is_zero(ch_inv[j])) allinv = Fr::mul(allinv, ch_inv[j]);
        VerifyInputs vin;
        for (size_t j = 0; j < lg_n; j++) { ch[j] = Fr::sqr(ch[j]); ch_inv[j] = Fr::sqr(ch_inv[j]); vin.usq[j] = ch[j]; }   // :292-296
        const fe& a = proof.a;
        const fe& b = proof.b;
        fe y_inv = Fr::inv(y);                                                                  // :473
        // device part: y^-i, g_scalars, h_scalars, delta
        DevBuf* wb[] = {&ctx->p_wL, &ctx->p_wR, &ctx->p_wO};
        for (auto* bf : wb) BP_CUDA_TRY(ctx, bf->reserve((n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->p_yinv.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->v_g.reserve((padded_n + 1) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->v_h.reserve((padded_n + 1) * sizeof(fe)));
        out.g = ctx->v_g.as<fe>();
        out.h = ctx->v_h.as<fe>();
        const int VB = 296;
        BP_CUDA_TRY(ctx, ctx->ipa_parts.reserve((size_t)(VB + 8) * sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->small.reserve(4096));
        if (int rc = D::pow_vec(ctx, y_inv, ctx->p_yinv.as<fe>(), padded_n)) return rc;          // :474-476
        vin.wL = ctx->p_wL.as<fe>(); vin.wR = ctx->p_wR.as<fe>(); vin.wO = ctx->p_wO.as<fe>(); vin.yinvpow = ctx->p_yinv.as<fe>();
        vin.allinv = allinv; vin.x = x; vin.a = a; vin.b = b; vin.u = u; vin.lg_n = (int)lg_n;
        int blocks = (int)((padded_n + 127) / 128);
        if (blocks > VB) blocks = VB;
        r1cs_verify_scalars_kernel<C><<<blocks, 128, 0, st>>>(vin, n, padded_n, n1, out.g, out.h, ctx->ipa_parts.as<fe>());
        BP_LAUNCH_CHECK(ctx);
        fe* d_delta = ctx->small.as<fe>() + 40;
        vec_reduce_partials_kernel<C, 1><<<1, 128, 0, st>>>(ctx->ipa_parts.as<fe>(), blocks, d_delta);
        BP_LAUNCH_CHECK(ctx);
        fe delta;
        if (int rc = D::download(ctx, &delta, d_delta, sizeof(fe))) return rc;
        if (ctx->world > 1) {
            // delta = <y^-n o w_R, w_L> runs over all multipliers; every rank has summed the indices it holds
            std::vector<uint8_t> all((size_t)ctx->world * sizeof(fe));
            if (int rc = ctx_allgather(ctx, (const uint8_t*)&delta, sizeof(fe), all.data())) return rc;
            delta = Fr::zero();
            for (int r2 = 0; r2 < ctx->world; r2++) {
                fe part;
                memcpy(&part, &all[(size_t)r2 * sizeof(fe)], sizeof(fe));
                delta = Fr::add(delta, part);
            }
        }
        fe xx = Fr::sqr(x), rxx = Fr::mul(r, xx), xxx = Fr::mul(x, xx);
        out.padded_n = padded_n;
        // scalars[0], scalars[1]  (:529-530)
        out.head[0] = Fr::add(Fr::mul(w, Fr::sub(proof.t_x, Fr::mul(a, b))), Fr::mul(r, Fr::sub(Fr::mul(xx, Fr::add(wc, delta)), proof.t_x)));
        out.head[1] = Fr::sub(Fr::neg(proof.e_blinding), Fr::mul(r, proof.t_x_blinding));
        out.tail.clear();
        out.tail.push_back(x); out.tail.push_back(xx); out.tail.push_back(xxx);                 // :533
        out.tail.push_back(Fr::mul(u, x)); out.tail.push_back(Fr::mul(u, xx)); out.tail.push_back(Fr::mul(u, xxx));
        for (auto& wVi : wV) out.tail.push_back(Fr::mul(wVi, rxx));                             // :534-536
        out.tail.push_back(Fr::mul(r, x)); out.tail.push_back(Fr::mul(rxx, x)); out.tail.push_back(Fr::mul(rxx, xx));   // T_scalars :526
        out.tail.push_back(Fr::mul(rxx, xxx)); out.tail.push_back(Fr::mul(Fr::mul(rxx, xx), xx));
        for (auto& s : ch) out.tail.push_back(s);                                               // u_sq      :538
        for (auto& s : ch_inv) out.tail.push_back(s);                                           // u_inv_sq  :539
        return BP_OK;
    }

    void tail_points(const ProofT<C>& proof, std::vector<affine>& pts) const {                  // verifier.rs:579-588
        const affine* p6[6] = {&proof.A_I1, &proof.A_O1, &proof.S1, &proof.A_I2, &proof.A_O2, &proof.S2};
        for (auto p : p6) pts.push_back(*p);
        for (auto& p : V) pts.push_back(p);
        const affine* p5[5] = {&proof.T_1, &proof.T_3, &proof.T_4, &proof.T_5, &proof.T_6};
        for (auto p : p5) pts.push_back(*p);
        for (auto& p : proof.L_vec) pts.push_back(p);
        for (auto& p : proof.R_vec) pts.push_back(p);
    }

    // verify_and_return_transcript (verifier.rs:559-600): one mega-MSM, accept iff identity
    int verify(const ProofT<C>& proof, const GensDev& gens) {
        StageTimer tm(ctx);
        Scalars sc;
        if (int rc = verification_scalars(proof, gens, sc)) return rc;
        tm.lap(ST_VSCALARS);
        std::vector<affine> pts;
        tail_points(proof, pts);
        int rc = mega_check(ctx, gens, sc.head, sc.g, sc.h, sc.padded_n, pts, sc.tail);
        tm.lap(ST_VMSM);
        return rc;
    }

    static int mega_check(bp_ctx* ctx, const GensDev& gens, const fe head[2], const fe* d_g, const fe* d_h, size_t np,
                          const std::vector<affine>& pts, const std::vector<fe>& tail) {
        affine sum;
        int ident = 0;
        if (int rc = mega_msm(ctx, gens, head, d_g, d_h, np, pts, tail, sum, ident)) return rc;
        return ident ? BP_OK : BP_ERR_VERIFY;                                                   // :595-597
    }

    static int mega_msm(bp_ctx* ctx, const GensDev& gens, const fe head[2], const fe* d_g, const fe* d_h, size_t np,
                        const std::vector<affine>& pts, const std::vector<fe>& tail, affine& sum, int& is_identity) {
        if (pts.size() != tail.size()) return BP_ERR_LEN;
        BP_CUDA_TRY(ctx, ctx->v_pts.reserve((pts.size() + 1) * sizeof(affine)));
        BP_CUDA_TRY(ctx, ctx->v_sc.reserve((tail.size() + 4) * sizeof(fe)));
        fe* d_sc = ctx->v_sc.as<fe>();
        D::upload(ctx, d_sc, head, 2 * sizeof(fe));
        D::upload(ctx, d_sc + 2, tail.data(), tail.size() * sizeof(fe));
        if (int rc = D::upload(ctx, ctx->v_pts.p, pts.data(), pts.size() * sizeof(affine))) return rc;
        // sharded contexts: G/H terms on the rank that holds them, everything else once (rank 0)
        const bool lead = ctx->rank == 0;
        MsmJob job;
        if (lead) job.add(gens.pc.template as<affine>(), d_sc, 2, 0);
        gens.add_range(job, gens.G, d_g, 0, np, 0);
        gens.add_range(job, gens.H, d_h, 0, np, 0);
        if (lead && !pts.empty()) job.add(ctx->v_pts.as<affine>(), d_sc + 2, pts.size(), 0);
        uint8_t o[1][64];
        int id[1] = {0};
        if (int rc = msm_run_job_sharded<C>(ctx, job, 1, o, id)) return rc;
        memcpy(&sum, o[0], 64);
        is_identity = id[0];
        return BP_OK;
    }
};

// IPA challenges of many verifier transcripts in one launch (transcript_dev.cuh). heads[p] must come from vs_head;
// proofs whose head failed are skipped; an identity L_j / R_j sets head_rc[p] = BP_ERR_VERIFY.
template <class C>
int device_ipa_challenges(bp_ctx* ctx, std::vector<VerifierT<C>*>& verifiers, std::vector<const ProofT<C>*>& proofs,
                          std::vector<typename VerifierT<C>::Head>& heads, std::vector<int>& head_rc) {
    const size_t k = verifiers.size();
    std::vector<DevTranscriptIn> in(k);
    std::vector<affine> pts;
    for (size_t p = 0; p < k; p++) {
        DevTranscriptIn& d = in[p];
        memset(&d, 0, sizeof(d));
        if (head_rc[p]) continue;                                  // lg_n = 0: the thread only draws an (unused) r
        const Strobe128& s = verifiers[p]->transcript->strobe;
        memcpy(d.st, s.state, 200);
        d.pos = s.pos; d.pos_begin = s.pos_begin;
        d.lg_n = (uint32_t)heads[p].lg_n;
        d.padded_n = heads[p].padded_n;
        if (pts.size() > 0xFFFFFFFFull - 64) return BP_ERR_LEN;
        d.pt_off = (uint32_t)pts.size();
        pts.insert(pts.end(), proofs[p]->L_vec.begin(), proofs[p]->L_vec.end());
        pts.insert(pts.end(), proofs[p]->R_vec.begin(), proofs[p]->R_vec.end());
    }
    cudaStream_t st = ctx->stream;
    const size_t out_fe = k * 32 * 2 + k;                          // u, u^-1 (32 slots per proof), r
    BP_CUDA_TRY(ctx, ctx->tr_in.reserve(k * sizeof(DevTranscriptIn)));
    BP_CUDA_TRY(ctx, ctx->tr_pts.reserve((pts.size() + 1) * sizeof(affine)));
    BP_CUDA_TRY(ctx, ctx->tr_out.reserve(out_fe * sizeof(fe) + k));
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tr_in.p, in.data(), k * sizeof(DevTranscriptIn), cudaMemcpyHostToDevice, st));
    if (!pts.empty()) BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tr_pts.p, pts.data(), pts.size() * sizeof(affine), cudaMemcpyHostToDevice, st));
    fe* d_u = ctx->tr_out.as<fe>();
    fe* d_ui = d_u + k * 32;
    fe* d_r = d_ui + k * 32;
    uint8_t* d_status = reinterpret_cast<uint8_t*>(d_r + k);
    verifier_ipa_challenges_kernel<C><<<(unsigned)((k + 63) / 64), 64, 0, st>>>(ctx->tr_in.as<DevTranscriptIn>(), ctx->tr_pts.as<affine>(), k, d_u, d_ui, d_r, d_status);
    BP_LAUNCH_CHECK(ctx);
    std::vector<uint8_t> back(out_fe * sizeof(fe) + k);
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(back.data(), ctx->tr_out.p, back.size(), cudaMemcpyDeviceToHost, st));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const fe* h_u = reinterpret_cast<const fe*>(back.data());
    const fe* h_ui = h_u + k * 32;
    const fe* h_r = h_ui + k * 32;
    const uint8_t* h_status = reinterpret_cast<const uint8_t*>(h_r + k);
    for (size_t p = 0; p < k; p++) {
        if (head_rc[p]) continue;
        if (h_status[p]) { head_rc[p] = BP_ERR_VERIFY; continue; }  // validate_and_append_point (transcript.rs:81-93)
        const size_t lg = heads[p].lg_n;
        heads[p].ch.assign(h_u + p * 32, h_u + p * 32 + lg);
        heads[p].ch_inv.assign(h_ui + p * 32, h_ui + p * 32 + lg);
        heads[p].r = h_r[p];
    }
    return BP_OK;
}

// ---- batch_verify (src/r1cs/verifier.rs:604-691) -------------------------------------------------
template <class C>
int batch_verify_t(bp_ctx* ctx, Rng* prng, const fe* alphas, std::vector<VerifierT<C>*>& verifiers, std::vector<const ProofT<C>*>& proofs,
                   const GensDev& gens, affine* partial_out = nullptr, int* partial_identity = nullptr) {
    using Fr = HostFp<typename C::Fr>;
    using HC = HostCurve<C>;
    size_t k = verifiers.size();
    if (proofs.size() != k) return BP_ERR_LEN;
    // The shared G/H accumulators need max_n_padded up front (:613-627); it only depends on the
    // constraint systems, not on the proofs' scalars, and the randomised phase can still add
    // multipliers, so a first pass computes every proof's scalars exactly like the reference and
    // the accumulators are grown on demand (a smaller proof simply leaves the tail untouched).
    fe head[2] = {Fr::zero(), Fr::zero()};
    std::vector<affine> pts;
    std::vector<fe> tail;
    size_t max_n = 0, acc_n = 0;
    // Large batches derive the inner-product challenges of all proofs in one launch (transcript_dev.cuh): every proof's
    // host transcript runs through `w` first, the device appends the (L_j, R_j) and draws u_j, u_j^-1 and r for all of
    // them, and the per-proof tails follow. Errors are still reported in proof order, like the reference's loop.
    const bool dev_tr = ctx->dev_transcript_min > 0 && k >= (size_t)ctx->dev_transcript_min;
    std::vector<typename VerifierT<C>::Head> heads(dev_tr ? k : 0);
    std::vector<int> head_rc(dev_tr ? k : 0, BP_OK);
    if (dev_tr) {
        for (size_t p = 0; p < k; p++) head_rc[p] = verifiers[p]->vs_head(*proofs[p], gens, heads[p]);
        if (int rc = device_ipa_challenges<C>(ctx, verifiers, proofs, heads, head_rc)) return rc;
    }
    for (size_t p = 0; p < k; p++) {
        typename VerifierT<C>::Scalars sc;
        if (dev_tr) {
            if (head_rc[p]) return head_rc[p];
            if (int rc = verifiers[p]->vs_tail(*proofs[p], gens, heads[p], sc, k > 1)) return rc;
        } else if (int rc = verifiers[p]->verification_scalars(*proofs[p], gens, sc, k > 1)) return rc;           // :619
        size_t np = sc.padded_n;
        if (np > max_n) max_n = np;
        if (np > acc_n) {
            // grow the accumulators, keeping what is already summed, zero-extending (:631-633)
            size_t want = np;
            DevBuf ng, nh;
            struct Owner { DevBuf *a, *b; bool keep = false; ~Owner() { if (!keep) { a->release(); b->release(); } } } own{&ng, &nh};   // not leaked on the error paths
            BP_CUDA_TRY(ctx, ng.reserve((want + 1) * sizeof(fe)));
            BP_CUDA_TRY(ctx, nh.reserve((want + 1) * sizeof(fe)));
            BP_CUDA_TRY(ctx, cudaMemsetAsync(ng.p, 0, want * sizeof(fe), ctx->stream));
            BP_CUDA_TRY(ctx, cudaMemsetAsync(nh.p, 0, want * sizeof(fe), ctx->stream));
            if (acc_n) {
                BP_CUDA_TRY(ctx, cudaMemcpyAsync(ng.p, ctx->v_accg.p, acc_n * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
                BP_CUDA_TRY(ctx, cudaMemcpyAsync(nh.p, ctx->v_acch.p, acc_n * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->v_accg.release(); ctx->v_acch.release();
            ctx->v_accg = ng; ctx->v_acch = nh;
            own.keep = true;
            acc_n = want;
        }
        fe alpha = alphas ? alphas[p] : HC::scalar_rand(*prng);                                 // :649 (same draw order)
        head[0] = Fr::add(head[0], Fr::mul(alpha, sc.head[0]));                                 // :652-653
        head[1] = Fr::add(head[1], Fr::mul(alpha, sc.head[1]));
        unsigned grid = (unsigned)((np + 255) / 256);
        vec_scale_accum_kernel<C><<<grid, 256, 0, ctx->stream>>>(sc.g, alpha, ctx->v_accg.as<fe>(), np);   // :655-664
        BP_LAUNCH_CHECK(ctx);
        vec_scale_accum_kernel<C><<<grid, 256, 0, ctx->stream>>>(sc.h, alpha, ctx->v_acch.as<fe>(), np);
        BP_LAUNCH_CHECK(ctx);
        for (auto& s : sc.tail) tail.push_back(Fr::mul(alpha, s));                              // :666-668
        verifiers[p]->tail_points(*proofs[p], pts);                                             // :669-682
    }
    if (k == 0) {
        BP_CUDA_TRY(ctx, ctx->v_accg.reserve(sizeof(fe)));
        BP_CUDA_TRY(ctx, ctx->v_acch.reserve(sizeof(fe)));
    }
    if (partial_out) {
        // multi-GPU batch verification: this rank's share of the final MSM; the caller
        // all-gathers the partial points and accepts iff their sum is the identity
        return VerifierT<C>::mega_msm(ctx, gens, head, ctx->v_accg.as<fe>(), ctx->v_acch.as<fe>(), max_n, pts, tail, *partial_out, *partial_identity);
    }
    return VerifierT<C>::mega_check(ctx, gens, head, ctx->v_accg.as<fe>(), ctx->v_acch.as<fe>(), max_n, pts, tail);   // :685-690
}

}  // namespace bp
