// pnp_precond.cu -- SSOR preconditioner (placeholder until the multicolour sweep lands).
#include "pnp_common.cuh"
namespace pnp {
void ssor_setup(Ctx&, Solver&, const Matrix&) {
  PNP_REQUIRE(false, PNP_E_ARG, "SSOR preconditioner not implemented yet");
}
void ssor_apply(Ctx&, Solver&, const Matrix&, const double*, double*) {}
} // namespace pnp
