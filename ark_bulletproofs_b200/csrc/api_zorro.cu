// zorro instantiation of the prover / verifier / IPA host layer and kernels.
#include "api_impl.cuh"
namespace bp { const CurveApi* curve_api_zorro() { return ApiImpl<Zorro>::table(); } }
