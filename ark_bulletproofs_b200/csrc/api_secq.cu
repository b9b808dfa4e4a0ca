// secq256k1 instantiation of the prover / verifier / IPA host layer and kernels.
#include "api_impl.cuh"
namespace bp { const CurveApi* curve_api_secq() { return ApiImpl<Secq256k1>::table(); } }
