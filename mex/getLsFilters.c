/* MEX drop-in for lib/getLsFilters.m:1-2, binding emagls_design_ls().
 * [wLsL, wLsR] = getLsFilters(hL, hR, hrirGridAziRad, hrirGridZenRad, order, shDefinition, shFunction)
 * Build: mex -R2018a -I../include getLsFilters.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 5) mexErrMsgIdAndTxt("eMagLS:nargin", "getLsFilters needs at least 5 arguments");
  emx_require_default_handle(nrhs, prhs, 6, "getSH");
  emagls_config cfg; emagls_config_default(&cfg);
  cfg.basis = emx_basis(nrhs, prhs, 5);
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]);
  const int order = (int)mxGetScalar(prhs[4]), nsh = (order + 1) * (order + 1);
  mxArray* wL = emx_out(T, nsh, cfg.basis); mxArray* wR = emx_out(T, nsh, cfg.basis);
  emx_check(emagls_design_ls(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                             mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), order, emx_ptr(wL), emx_ptr(wR)));
  emx_return2(nlhs, plhs, wL, wR);
}
