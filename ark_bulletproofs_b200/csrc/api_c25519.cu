// curve25519 instantiation of the prover / verifier / IPA host layer and kernels.
#include "api_impl.cuh"
namespace bp { const CurveApi* curve_api_c25519() { return ApiImpl<Curve25519>::table(); } }
