// mpcv_params.h — mpcv_spec (C ABI) -> Params (kernel argument), with IPOPT 3.12 defaults for
// unset options.  Option names follow the `opts['ipopt']` dict of the scripts
// (Casadi/single_shooting_v1.py:121-129).
#pragma once

#include "mpcv_models.cuh"

namespace mpcv {

inline Params params_from_spec(const mpcv_spec& s) {
  Params P;
  P.N = s.N;
  P.M = s.M > 0 ? s.M : 1;
  P.ntu = s.ntu;
  P.T = s.T;
  for (int i = 0; i < 4; ++i) { P.Q[i] = s.Q[i]; P.extra[i] = s.extra[i]; }
  P.R[0] = s.R[0]; P.R[1] = s.R[1];
  P.R1 = s.R1;
  P.tol = s.tol > 0 ? s.tol : 1e-8;
  P.max_iter = s.max_iter > 0 ? s.max_iter : 3000;
  P.max_soc = s.max_soc >= 0 ? s.max_soc : 4;
  P.mu_init = s.mu_init > 0 ? s.mu_init : 0.1;
  P.bound_push = s.bound_push > 0 ? s.bound_push : 1e-2;
  P.bound_frac = s.bound_frac > 0 ? s.bound_frac : 1e-2;
  P.bound_relax = s.bound_relax_factor >= 0 ? s.bound_relax_factor : 1e-8;
  P.scal_max_grad = s.nlp_scaling_max_gradient > 0 ? s.nlp_scaling_max_gradient : 100.0;
  P.dual_inf_tol = s.dual_inf_tol > 0 ? s.dual_inf_tol : 1.0;
  P.constr_viol_tol = s.constr_viol_tol > 0 ? s.constr_viol_tol : 1e-4;
  P.compl_inf_tol = s.compl_inf_tol > 0 ? s.compl_inf_tol : 1e-4;
  P.acceptable_tol = s.acceptable_tol > 0 ? s.acceptable_tol : 1e-6;
  P.acceptable_iter = s.acceptable_iter != 0 ? s.acceptable_iter : 15;
  P.acceptable_obj_change_tol = s.acceptable_obj_change_tol > 0 ? s.acceptable_obj_change_tol : 1e20;
  P.diag = nullptr;
  return P;
}

}  // namespace mpcv
