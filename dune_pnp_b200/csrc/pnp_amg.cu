// pnp_amg.cu -- aggregation AMG preconditioner (placeholder).
#include "pnp_common.cuh"
namespace pnp {
struct Amg {};
void amg_setup(Ctx&, Solver&, const Matrix&) {
  PNP_REQUIRE(false, PNP_E_ARG, "AMG preconditioner not implemented yet");
}
void amg_apply(Ctx&, Solver&, const Matrix&, const double*, double*) {}
} // namespace pnp
