// Per-curve implementation of the handle-based C ABI (include/bp_b200.h), instantiated once per
// curve in api_<curve>.cu and reached from capi.cu through a table of function pointers.
#pragma once
#include "api_types.hpp"
#include "r1cs.cuh"
#include "gens_kernels.cuh"

namespace bp {

template <class C>
struct ApiImpl {
    using HC = HostCurve<C>;
    using Fr = HostFp<typename C::Fr>;

    static fe ld(const uint8_t* p) { fe r; memcpy(r.v, p, 32); return r; }
    static affine ldp(const uint8_t* p) { affine r; memcpy(&r, p, 64); return r; }

    // ---- generators -----------------------------------------------------------------------
    static int gens_upload(bp_ctx* ctx, const affine& B, const affine& Bb, const affine* G, const affine* H, size_t cap, GensDev** out) {
        std::unique_ptr<GensDev> g(new GensDev());
        g->ctx = ctx;
        g->capacity = cap;
        g->B = B;
        g->B_blinding = Bb;
        g->rank = ctx->rank;
        g->world = ctx->world;
        std::vector<affine> lg, lh;
        size_t lcap = cap;
        if (ctx->world > 1) {
            // keep this rank's cyclic shard only: generator i lives on rank i mod world
            size_t lo;
            g->slice(0, cap, lo, lcap);
            lg.resize(lcap); lh.resize(lcap);
            for (size_t j = 0; j < lcap; j++) { lg[j] = G[j * ctx->world + ctx->rank]; lh[j] = H[j * ctx->world + ctx->rank]; }
            G = lg.data(); H = lh.data();
        }
        BP_CUDA_TRY(ctx, g->G.reserve((lcap + 1) * sizeof(affine)));
        BP_CUDA_TRY(ctx, g->H.reserve((lcap + 1) * sizeof(affine)));
        BP_CUDA_TRY(ctx, g->pc.reserve(2 * sizeof(affine)));
        affine pc[2] = {B, Bb};
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(g->pc.p, pc, sizeof(pc), cudaMemcpyHostToDevice, ctx->stream));
        if (lcap) {
            BP_CUDA_TRY(ctx, cudaMemcpyAsync(g->G.p, G, lcap * sizeof(affine), cudaMemcpyHostToDevice, ctx->stream));
            BP_CUDA_TRY(ctx, cudaMemcpyAsync(g->H.p, H, lcap * sizeof(affine), cudaMemcpyHostToDevice, ctx->stream));
        }
        BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        *out = g.release();
        return BP_OK;
    }
    static int gens_generate_host(size_t cap, uint8_t* G, uint8_t* H, uint8_t* B, uint8_t* Bb) {
        affine b, bb;
        GensHost<C>::pedersen_default(b, bb);
        if (B) memcpy(B, &b, 64);
        if (Bb) memcpy(Bb, &bb, 64);
        if (cap && G && H) GensHost<C>::bulletproof_gens(cap, reinterpret_cast<affine*>(G), reinterpret_cast<affine*>(H));
        return BP_OK;
    }
    static int gens_create(bp_ctx* ctx, size_t cap, GensDev** out) {
        affine b, bb;
        GensHost<C>::pedersen_default(b, bb);
        // The chains are generated on the device (gens_kernels.cuh): every Affine::rand attempt of secq256k1 and
        // curve25519 reads nine keystream words; zorro's x draw rejects, so the host walks the keystream and the device
        // evaluates the accepted draws. Tiny capacities and the astronomically unlikely irregular stream use the host generator.
        {
            if (ctx->gens_on_device && cap >= 256) {
                std::unique_ptr<GensDev> g(new GensDev());
                g->ctx = ctx; g->capacity = cap; g->B = b; g->B_blinding = bb; g->rank = ctx->rank; g->world = ctx->world;
                size_t lo, lcap;
                g->slice(0, cap, lo, lcap);
                BP_CUDA_TRY(ctx, g->G.reserve((lcap + 1) * sizeof(affine)));
                BP_CUDA_TRY(ctx, g->H.reserve((lcap + 1) * sizeof(affine)));
                BP_CUDA_TRY(ctx, g->pc.reserve(2 * sizeof(affine)));
                affine pc[2] = {b, bb};
                BP_CUDA_TRY(ctx, cudaMemcpyAsync(g->pc.p, pc, sizeof(pc), cudaMemcpyHostToDevice, ctx->stream));
                const auto& ts = HC::ts_params();
                SqrtParams sp;
                memcpy(sp.t, ts.t, 32); memcpy(sp.t1h, ts.t1h, 32); sp.z = ts.z; sp.s = ts.s;
                uint8_t label[5] = {'G', 0, 0, 0, 0}, seed[32];                                   // generators.rs:196-221, party 0
                GensHost<C>::chain_seed(label, 5, seed);
                int rc = gens_chain_device<C>(ctx, seed, sp, cap, ctx->rank, ctx->world, g->G.template as<affine>());
                if (rc == BP_OK) {
                    label[0] = 'H';
                    GensHost<C>::chain_seed(label, 5, seed);
                    rc = gens_chain_device<C>(ctx, seed, sp, cap, ctx->rank, ctx->world, g->H.template as<affine>());
                }
                if (rc == BP_OK) { *out = g.release(); return BP_OK; }
                if (rc != BP_ERR_UNSUPPORTED) return rc;
            }
        }
        std::vector<affine> G(cap), H(cap);
        GensHost<C>::bulletproof_gens(cap, G.data(), H.data());
        return gens_upload(ctx, b, bb, G.data(), H.data(), cap, out);
    }
    static int gens_from_points(bp_ctx* ctx, const uint8_t* B, const uint8_t* Bb, const uint8_t* G, const uint8_t* H, size_t cap, GensDev** out) {
        if (!HC::point_valid(ldp(B)) || !HC::point_valid(ldp(Bb))) return BP_ERR_FORMAT;
        GensDev* g = nullptr;
        if (int rc = gens_upload(ctx, ldp(B), ldp(Bb), reinterpret_cast<const affine*>(G), reinterpret_cast<const affine*>(H), cap, &g)) return rc;
        // the vectors are validated where they now live (canonical, on curve, prime-order subgroup), one thread per point
        size_t lo = 0, lcap = cap;
        if (ctx->world > 1) g->slice(0, cap, lo, lcap);
        int rc = points_validate_device<C>(ctx, g->G.template as<affine>(), lcap);
        if (rc == BP_OK) rc = points_validate_device<C>(ctx, g->H.template as<affine>(), lcap);
        if (rc != BP_OK) { delete g; return rc; }
        *out = g;
        return BP_OK;
    }
    static int pedersen_commit(const GensDev* g, const uint8_t* v, const uint8_t* blind, uint8_t* out) {
        affine r = HC::add(HC::mul(g->B, ld(v)), HC::mul(g->B_blinding, ld(blind)));
        memcpy(out, &r, 64);
        return BP_OK;
    }

    // ---- scalars / points ---------------------------------------------------------------------
    static int challenge_scalar(Transcript* t, const char* label, uint8_t* out) {
        fe s = TP<C>::challenge_scalar(*t, label);
        memcpy(out, s.v, 32);
        return BP_OK;
    }
    static int rng_scalar(Rng* rng, uint8_t* out) {
        fe s = HC::scalar_rand(*rng);
        memcpy(out, s.v, 32);
        return BP_OK;
    }
    static int rng_scalars(Rng* rng, size_t n, uint8_t* out) {
        HC::scalar_rand_bulk(*rng, reinterpret_cast<fe*>(out), n);
        return BP_OK;
    }
    static int scalar_to_bytes(const uint8_t* mont, uint8_t* out) { HC::scalar_to_bytes(ld(mont), out); return BP_OK; }
    static int scalar_from_bytes(const uint8_t* in, uint8_t* mont) {
        fe s;
        if (!HC::scalar_from_bytes(in, s)) return BP_ERR_FORMAT;
        memcpy(mont, s.v, 32);
        return BP_OK;
    }
    static int point_compress(const uint8_t* xy, uint8_t* out) { HC::point_compressed(ldp(xy), out); return BP_OK; }
    static int point_uncompressed(const uint8_t* xy, uint8_t* out) { HC::point_uncompressed(ldp(xy), out); return BP_OK; }
    static int point_decompress(const uint8_t* in, uint8_t* xy) {
        affine p;
        if (!HC::point_from_compressed(in, p)) return BP_ERR_FORMAT;
        memcpy(xy, &p, 64);
        return BP_OK;
    }

    // ---- prover / verifier / proof ------------------------------------------------------------------
    static void* prover_new(bp_ctx* ctx, const GensDev* g, Transcript* t) { return new ProverT<C>(ctx, g, t); }
    static void prover_free(void* p) { delete static_cast<ProverT<C>*>(p); }
    static ConstraintSystemBase* prover_cs(void* p) { return static_cast<ProverT<C>*>(p); }
    static int prover_commit(void* p, const uint8_t* v, const uint8_t* blind, uint8_t* out_V, Variable* var) {
        affine V;
        int rc = static_cast<ProverT<C>*>(p)->commit(ld(v), ld(blind), V, *var);
        memcpy(out_V, &V, 64);
        return rc;
    }
    static int prover_commit_batch(void* p, const uint8_t* v, const uint8_t* blind, size_t m, uint8_t* out_V, Variable* vars) {
        return static_cast<ProverT<C>*>(p)->commit_batch(reinterpret_cast<const fe*>(v), reinterpret_cast<const fe*>(blind), m,
                                                        reinterpret_cast<affine*>(out_V), vars);
    }
    static int prover_prove(void* p, Rng* rng, void** out_proof) {
        std::unique_ptr<ProofT<C>> pr(new ProofT<C>());
        int rc = static_cast<ProverT<C>*>(p)->prove(*rng, *pr);
        if (rc == BP_OK) *out_proof = pr.release();
        return rc;
    }
    static void* verifier_new(bp_ctx* ctx, Transcript* t) { return new VerifierT<C>(ctx, t); }
    static void verifier_free(void* p) { delete static_cast<VerifierT<C>*>(p); }
    static ConstraintSystemBase* verifier_cs(void* p) { return static_cast<VerifierT<C>*>(p); }
    static int verifier_commit(void* p, const uint8_t* V, Variable* var) {
        const affine c = ldp(V);
        if (!HC::point_valid(c)) return BP_ERR_FORMAT;           // attacker-chosen public input: never reaches the transcript or the MSM unchecked
        return static_cast<VerifierT<C>*>(p)->commit(c, *var);
    }
    static int verifier_verify(void* p, const void* proof, const GensDev* g) {
        return static_cast<VerifierT<C>*>(p)->verify(*static_cast<const ProofT<C>*>(proof), *g);
    }
    static int batch_verify(bp_ctx* ctx, Rng* rng, void** verifiers, const void** proofs, size_t n, const GensDev* g) {
        std::vector<VerifierT<C>*> vs(n);
        std::vector<const ProofT<C>*> ps(n);
        for (size_t i = 0; i < n; i++) { vs[i] = static_cast<VerifierT<C>*>(verifiers[i]); ps[i] = static_cast<const ProofT<C>*>(proofs[i]); }
        return batch_verify_t<C>(ctx, rng, nullptr, vs, ps, *g);
    }
    static int batch_verify_partial(bp_ctx* ctx, const uint8_t* alphas, void** verifiers, const void** proofs, size_t n, const GensDev* g,
                                    uint8_t* out_xy, int* out_identity) {
        std::vector<VerifierT<C>*> vs(n);
        std::vector<const ProofT<C>*> ps(n);
        std::vector<fe> al(n);
        for (size_t i = 0; i < n; i++) {
            vs[i] = static_cast<VerifierT<C>*>(verifiers[i]);
            ps[i] = static_cast<const ProofT<C>*>(proofs[i]);
            al[i] = ld(alphas + 32 * i);
        }
        affine sum;
        int ident = 0;
        int rc = batch_verify_t<C>(ctx, nullptr, al.data(), vs, ps, *g, &sum, &ident);
        if (rc) return rc;
        memcpy(out_xy, &sum, 64);
        *out_identity = ident;
        return BP_OK;
    }
    static void proof_free(void* p) { delete static_cast<ProofT<C>*>(p); }
    static int proof_to_bytes(const void* p, std::vector<uint8_t>& out) { out = static_cast<const ProofT<C>*>(p)->to_bytes(); return BP_OK; }
    static int proof_from_bytes(const uint8_t* d, size_t len, void** out) {
        std::unique_ptr<ProofT<C>> pr(new ProofT<C>());
        int rc = ProofT<C>::from_bytes(d, len, *pr);
        if (rc == BP_OK) *out = pr.release();
        return rc;
    }
    // Many proofs at once (batch-verification ingest): the structure and the scalars are parsed on the host, all
    // compressed points of all proofs are decompressed and validated in one GPU launch (gens_kernels.cuh). status[i] is
    // BP_OK or BP_ERR_FORMAT exactly as from_bytes would decide; out[i] is NULL for rejected proofs.
    static int proofs_from_bytes_batch(bp_ctx* ctx, const uint8_t* const* bufs, const size_t* lens, size_t n, void** out, int* status) {
        std::vector<std::unique_ptr<ProofT<C>>> prs(n);
        if constexpr (C::KIND != 0) {
            for (size_t i = 0; i < n; i++) {       // twisted Edwards: subgroup check per point; host path
                prs[i].reset(new ProofT<C>());
                status[i] = ProofT<C>::from_bytes(bufs[i], lens[i], *prs[i]);
                out[i] = status[i] == BP_OK ? prs[i].release() : nullptr;
            }
            return BP_OK;
        } else {
            std::vector<std::pair<const uint8_t*, affine*>> defer;
            std::vector<size_t> first(n + 1, 0);
            for (size_t i = 0; i < n; i++) {
                prs[i].reset(new ProofT<C>());
                first[i] = defer.size();
                status[i] = ProofT<C>::from_bytes(bufs[i], lens[i], *prs[i], &defer);
                if (status[i] != BP_OK) defer.resize(first[i]);
            }
            first[n] = defer.size();
            const size_t np = defer.size();
            std::vector<uint8_t> comp(np * 33), ok(np);
            std::vector<affine> pts(np);
            for (size_t k = 0; k < np; k++) memcpy(&comp[k * 33], defer[k].first, 33);
            const auto& ts = HC::ts_params();
            SqrtParams sp;
            memcpy(sp.t, ts.t, 32); memcpy(sp.t1h, ts.t1h, 32); sp.z = ts.z; sp.s = ts.s;
            if (int rc = points_decompress_device<C>(ctx, comp.data(), np, sp, pts.data(), ok.data())) return rc;
            for (size_t i = 0; i < n; i++) {
                if (status[i] == BP_OK) {
                    size_t lo = first[i], hi = first[i + 1];
                    for (size_t k = lo; k < hi; k++) {
                        if (!ok[k]) { status[i] = BP_ERR_FORMAT; break; }
                        *defer[k].second = pts[k];
                    }
                }
                out[i] = status[i] == BP_OK ? prs[i].release() : nullptr;
            }
            return BP_OK;
        }
    }
    static void* proof_clone(const void* p) { return new ProofT<C>(*static_cast<const ProofT<C>*>(p)); }
    // field access for tamper tests: which = 0 t_x, 1 t_x_blinding, 2 e_blinding, 3 a, 4 b (scalars);
    // 10.. = points A_I1,A_O1,S1,A_I2,A_O2,S2,T_1,T_3,T_4,T_5,T_6 ; 100+j = L_j ; 200+j = R_j
    static int proof_field(void* p, int which, uint8_t* buf, int set) {
        ProofT<C>* pr = static_cast<ProofT<C>*>(p);
        fe* s = nullptr;
        affine* a = nullptr;
        affine* pts[11] = {&pr->A_I1, &pr->A_O1, &pr->S1, &pr->A_I2, &pr->A_O2, &pr->S2, &pr->T_1, &pr->T_3, &pr->T_4, &pr->T_5, &pr->T_6};
        if (which == 0) s = &pr->t_x; else if (which == 1) s = &pr->t_x_blinding; else if (which == 2) s = &pr->e_blinding;
        else if (which == 3) s = &pr->a; else if (which == 4) s = &pr->b;
        else if (which >= 10 && which < 21) a = pts[which - 10];
        else if (which >= 100 && which < 100 + (int)pr->L_vec.size()) a = &pr->L_vec[which - 100];
        else if (which >= 200 && which < 200 + (int)pr->R_vec.size()) a = &pr->R_vec[which - 200];
        else return BP_ERR_ARG;
        if (s) { if (set) memcpy(s->v, buf, 32); else memcpy(buf, s->v, 32); }
        else if (set) {
            affine np;
            memcpy(&np, buf, 64);
            if (!HC::point_valid(np)) return BP_ERR_FORMAT;
            *a = np;
        } else memcpy(buf, a, 64);
        return BP_OK;
    }
    static size_t proof_rounds(const void* p) { return static_cast<const ProofT<C>*>(p)->L_vec.size(); }

    // The device transcript on one transcript (test hook of transcript_dev.cuh): u_j, u_j^-1 for j < lg_n and r, from
    // the state of `t` (which is not advanced), the domain separator for `padded_n` and the points L, R.
    static int ipa_challenges_device(bp_ctx* ctx, const Transcript* t, uint64_t padded_n, const uint8_t* L, const uint8_t* R, size_t lg_n,
                                     uint8_t* out_u, uint8_t* out_uinv, uint8_t* out_r, int* status) {
        if (lg_n > 32) return BP_ERR_LEN;
        DevTranscriptIn in;
        memset(&in, 0, sizeof(in));
        memcpy(in.st, t->strobe.state, 200);
        in.pos = t->strobe.pos; in.pos_begin = t->strobe.pos_begin;
        in.lg_n = (uint32_t)lg_n; in.padded_n = padded_n; in.pt_off = 0;
        std::vector<affine> pts(2 * lg_n + 1);
        for (size_t j = 0; j < lg_n; j++) { pts[j] = ldp(L + 64 * j); pts[lg_n + j] = ldp(R + 64 * j); }
        cudaStream_t st = ctx->stream;
        BP_CUDA_TRY(ctx, ctx->tr_in.reserve(sizeof(in)));
        BP_CUDA_TRY(ctx, ctx->tr_pts.reserve(pts.size() * sizeof(affine)));
        BP_CUDA_TRY(ctx, ctx->tr_out.reserve(65 * sizeof(fe) + 16));
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tr_in.p, &in, sizeof(in), cudaMemcpyHostToDevice, st));
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tr_pts.p, pts.data(), pts.size() * sizeof(affine), cudaMemcpyHostToDevice, st));
        fe* d_u = ctx->tr_out.template as<fe>();
        verifier_ipa_challenges_kernel<C><<<1, 64, 0, st>>>(ctx->tr_in.template as<DevTranscriptIn>(), ctx->tr_pts.template as<affine>(), 1, d_u, d_u + 32, d_u + 64,
                                                           reinterpret_cast<uint8_t*>(d_u + 65));
        BP_LAUNCH_CHECK(ctx);
        std::vector<uint8_t> back(65 * sizeof(fe) + 16);
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(back.data(), ctx->tr_out.p, back.size(), cudaMemcpyDeviceToHost, st));
        BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        memcpy(out_u, back.data(), lg_n * 32);
        memcpy(out_uinv, back.data() + 32 * 32, lg_n * 32);
        memcpy(out_r, back.data() + 64 * 32, 32);
        *status = back[65 * 32];
        return BP_OK;
    }
    static int chain_circuit(ConstraintSystemBase* cs, const Variable* v0, size_t n, const uint8_t* ks, const uint8_t* x0) {
        fe x = Fr::zero();
        if (x0) x = ld(x0);
        const fe one = Fr::one(), neg1 = Fr::neg(Fr::one());
        Variable prev_o{0, 0};
        for (size_t i = 0; i < n; i++) {
            fe k = ld(ks + 32 * i);
            Variable o[3];
            if (int rc = cs->allocate_multiplier(x0 ? &x : nullptr, x0 ? &k : nullptr, o)) return rc;
            Variable v2[2] = {o[1], {VAR_ONE, 0}};
            fe c2[2] = {one, Fr::neg(k)};
            if (int rc = cs->constrain(v2, c2, 2)) return rc;                 // R_i - k_i
            Variable v3[2] = {o[0], i == 0 ? *v0 : prev_o};
            fe c3[2] = {one, neg1};
            if (int rc = cs->constrain(v3, c3, 2)) return rc;                 // L_i - (V | O_{i-1})
            prev_o = o[2];
            if (x0) x = Fr::mul(x, k);
        }
        return BP_OK;
    }

    // The reference's benchmark / test gadget (benches/r1cs_secq256k1.rs:35-75, tests/r1cs_secq256k1.rs:16-56): y is a
    // permutation of x iff prod (x_i - z) == prod (y_i - z) for a random z; 2(k-1) phase-2 multipliers.
    static int shuffle_gadget(ConstraintSystemBase* cs, const Variable* x, const Variable* y, size_t k) {
        const fe one = Fr::one(), neg1 = Fr::neg(Fr::one());
        if (k == 0) return BP_ERR_ARG;
        if (k == 1) {
            Variable v[2] = {y[0], x[0]};
            fe c[2] = {one, neg1};
            return cs->constrain(v, c, 2);
        }
        std::vector<Variable> xs(x, x + k), ys(y, y + k);
        return cs->specify_randomized_constraints([xs, ys, k, one, neg1](ConstraintSystemBase& c) -> int {
            fe z;
            if (int rc = c.challenge_scalar("shuffle challenge", &z)) return rc;
            const fe cz[2] = {one, Fr::neg(z)};
            auto product = [&](const std::vector<Variable>& v, Variable& first) -> int {
                Variable o[3];
                Variable lv[2] = {v[k - 1], {VAR_ONE, 0}}, rv[2] = {v[k - 2], {VAR_ONE, 0}};
                if (int rc = c.multiply(lv, cz, 2, rv, cz, 2, o)) return rc;
                first = o[2];
                for (size_t i = k - 2; i-- > 0;) {
                    Variable l1[1] = {first}, r2[2] = {v[i], {VAR_ONE, 0}};
                    if (int rc = c.multiply(l1, &one, 1, r2, cz, 2, o)) return rc;
                    first = o[2];
                }
                return BP_OK;
            };
            Variable fx, fy;
            if (int rc = product(xs, fx)) return rc;
            if (int rc = product(ys, fy)) return rc;
            Variable v[2] = {fx, fy};
            fe cc[2] = {one, neg1};
            return c.constrain(v, cc, 2);
        });
    }

    // ---- InnerProductProof::create over host buffers --------------------------------------------
    static int ipa_create_host(bp_ctx* ctx, Transcript* t, const uint8_t* Q, const uint8_t* Gf, const uint8_t* Hf, const uint8_t* G,
                               const uint8_t* H, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out_L, uint8_t* out_R,
                               uint8_t* out_a, uint8_t* out_b) {
        if (n == 0 || (n & (n - 1))) return BP_ERR_POW2;
        DevBuf dG, dH, dGf, dHf, da, db;
        struct Guard { DevBuf* b[6]; ~Guard() { for (auto* x : b) x->release(); } } guard{{&dG, &dH, &dGf, &dHf, &da, &db}};
        BP_CUDA_TRY(ctx, dG.reserve(n * 64)); BP_CUDA_TRY(ctx, dH.reserve(n * 64));
        BP_CUDA_TRY(ctx, dGf.reserve(n * 32)); BP_CUDA_TRY(ctx, dHf.reserve(n * 32));
        BP_CUDA_TRY(ctx, da.reserve(n * 32)); BP_CUDA_TRY(ctx, db.reserve(n * 32));
        using D = Dev<C>;
        std::vector<affine> lg, lh;
        size_t nl = n;
        if (ctx->world > 1) {
            // multi-GPU context: the caller passes the full vectors on every rank; only this rank's cyclic shard
            // of G and H (index = rank mod world) is kept on the device
            if (n < (size_t)ctx->world || n % ctx->world) return BP_ERR_ARG;
            nl = n / ctx->world;
            lg.resize(nl); lh.resize(nl);
            for (size_t j = 0; j < nl; j++) {
                lg[j] = reinterpret_cast<const affine*>(G)[j * ctx->world + ctx->rank];
                lh[j] = reinterpret_cast<const affine*>(H)[j * ctx->world + ctx->rank];
            }
            G = reinterpret_cast<const uint8_t*>(lg.data());
            H = reinterpret_cast<const uint8_t*>(lh.data());
        }
        D::upload(ctx, dG.p, G, nl * 64); D::upload(ctx, dH.p, H, nl * 64); D::upload(ctx, dGf.p, Gf, n * 32);
        D::upload(ctx, dHf.p, Hf, n * 32); D::upload(ctx, da.p, a, n * 32);
        if (int rc = D::upload(ctx, db.p, b, n * 32)) return rc;
        std::vector<affine> L, R;
        fe ao, bo;
        int rc = ipa_create<C>(ctx, *t, ldp(Q), dGf.as<fe>(), dHf.as<fe>(), dG.as<affine>(), dH.as<affine>(), da.as<fe>(), db.as<fe>(), n, L, R, ao, bo);
        if (rc) return rc;
        if (!L.empty()) { memcpy(out_L, L.data(), L.size() * 64); memcpy(out_R, R.data(), R.size() * 64); }
        memcpy(out_a, ao.v, 32);
        memcpy(out_b, bo.v, 32);
        return BP_OK;
    }

    static int ipa_verify_host(bp_ctx* ctx, Transcript* t, size_t n, const uint8_t* Lp, const uint8_t* Rp, const uint8_t* a, const uint8_t* b,
                               const uint8_t* Gf, const uint8_t* Hf, const uint8_t* P, const uint8_t* Q, const uint8_t* G, const uint8_t* H) {
        if (n == 0 || (n & (n - 1))) return BP_ERR_POW2;
        size_t k = 0;
        while (((size_t)1 << k) < n) k++;
        DevBuf dG, dH, dGf, dHf;
        struct Guard { DevBuf* b[4]; ~Guard() { for (auto* x : b) x->release(); } } guard{{&dG, &dH, &dGf, &dHf}};
        BP_CUDA_TRY(ctx, dG.reserve(n * 64)); BP_CUDA_TRY(ctx, dH.reserve(n * 64));
        BP_CUDA_TRY(ctx, dGf.reserve(n * 32)); BP_CUDA_TRY(ctx, dHf.reserve(n * 32));
        using D = Dev<C>;
        D::upload(ctx, dG.p, G, n * 64); D::upload(ctx, dH.p, H, n * 64); D::upload(ctx, dGf.p, Gf, n * 32);
        if (int rc = D::upload(ctx, dHf.p, Hf, n * 32)) return rc;
        std::vector<affine> L(k), R(k);
        if (k) { memcpy(L.data(), Lp, k * 64); memcpy(R.data(), Rp, k * 64); }
        return ipa_verify<C>(ctx, *t, n, L, R, ld(a), ld(b), dGf.as<fe>(), dHf.as<fe>(), ldp(P), ldp(Q), dG.as<affine>(), dH.as<affine>());
    }

    static const CurveApi* table() {
        static const CurveApi api = {
            gens_generate_host, gens_create, gens_from_points, pedersen_commit, challenge_scalar, rng_scalar, scalar_to_bytes, scalar_from_bytes,
            point_compress, point_uncompressed, point_decompress, prover_new, prover_free, prover_cs, prover_commit, prover_commit_batch, prover_prove, verifier_new,
            verifier_free, verifier_cs, verifier_commit, verifier_verify, batch_verify, batch_verify_partial, proof_free, proof_to_bytes, proof_from_bytes, proof_clone,
            proof_field, proof_rounds, chain_circuit, ipa_create_host, ipa_verify_host, rng_scalars, shuffle_gadget, proofs_from_bytes_batch,
            ipa_challenges_device};
        return &api;
    }
};

}  // namespace bp
