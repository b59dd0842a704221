// Forwarding header: flux-based scheme objects (make_convection_upwind, make_diffusion_order2, `u - dt * scheme(u)`) live in
// b200_api.hpp (reference: schemes/fv.hpp -> schemes/fv/operators/*.hpp, flux_based/*.hpp).
#pragma once
#include "../b200_api.hpp"
