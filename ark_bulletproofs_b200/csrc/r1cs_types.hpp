// Types shared by the templated host layer (r1cs.cuh) and the C ABI glue (capi.cu).
#pragma once
#include <functional>
#include <vector>
#include "ctx.cuh"

namespace bp {

enum VarKind : uint32_t { VAR_COMMITTED = 0, VAR_MUL_LEFT = 1, VAR_MUL_RIGHT = 2, VAR_MUL_OUT = 3, VAR_ONE = 4 };
struct Variable { uint32_t kind; uint64_t idx; };

// ---- device-resident generators (BulletproofGens share 0 + PedersenGens) ------------------------
struct GensDev {
    bp_ctx* ctx = nullptr;
    size_t capacity = 0;      // global capacity (BulletproofGens::gens_capacity)
    // cyclic shard of a multi-GPU context (bp_ctx_set_collective): G[j], H[j] hold generator j*world + rank
    int rank = 0, world = 1;
    DevBuf G, H, pc;          // pc = [B, B_blinding]
    mutable DevBuf pc_table;  // fixed-base table of B and B_blinding (vec_kernels.cuh), built on the first batched commit
    mutable bool has_pc_table = false;
    affine B, B_blinding;
    ~GensDev() { G.release(); H.release(); pc.release(); pc_table.release(); }
    // generators [off, off+cnt) of the global numbering -> local entries [lo, hi); local j is global j*world + rank
    void slice(size_t off, size_t cnt, size_t& lo, size_t& hi) const {
        size_t end = off + cnt, r = (size_t)rank, w = (size_t)world;
        lo = off > r ? (off - r + w - 1) / w : 0;
        hi = end > r ? (end - r + w - 1) / w : 0;
    }
    // adds <scalars[0..cnt), generators[off..off+cnt)> to MSM m of the job; `scalars` is the replicated vector
    template <class Job> void add_range(Job& job, const DevBuf& gen, const fe* scalars, size_t off, size_t cnt, int m) const {
        size_t lo, hi;
        slice(off, cnt, lo, hi);
        if (hi > lo) job.add(gen.template as<affine>() + lo, scalars + (lo * world + rank - off), hi - lo, m, (uint32_t)world);
    }
};

// ---- abstract interfaces for the C ABI ------------------------------------------------------------
struct ProofBase {
    virtual ~ProofBase() {}
    virtual std::vector<uint8_t> to_bytes() const = 0;
};
struct ConstraintSystemBase {
    virtual ~ConstraintSystemBase() {}
    virtual int multiply(const Variable* lv, const fe* lc, size_t ln, const Variable* rv, const fe* rc, size_t rn, Variable out[3]) = 0;
    virtual int allocate(const fe* assignment, Variable* out) = 0;
    virtual int allocate_multiplier(const fe* l, const fe* r, Variable out[3]) = 0;
    virtual int constrain(const Variable* v, const fe* c, size_t n) = 0;
    virtual size_t multipliers_len() const = 0;
    virtual int challenge_scalar(const char* label, fe* out) = 0;
    virtual int specify_randomized_constraints(std::function<int(ConstraintSystemBase&)> cb) = 0;
};

}  // namespace bp
