/* MEX drop-in for lib/getEMagLsFilters.m:1-2, binding emagls_design_emagls().
 * [wMlsL, wMlsR] = getEMagLsFilters(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, micGridAziRad,
 *                                   micGridZenRad, order, fs, len, shDefinition, shFunction)
 * Build: mex -R2018a -I../include getEMagLsFilters.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 10) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLsFilters needs at least 10 arguments");
  emx_require_default_handle(nrhs, prhs, 11, "getSH");
  emagls_config cfg; emagls_config_default(&cfg);
  cfg.basis = emx_basis(nrhs, prhs, 10);
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]), M = (int)mxGetNumberOfElements(prhs[5]);
  const int order = (int)mxGetScalar(prhs[7]), len = (int)mxGetScalar(prhs[9]), nsh = (order + 1) * (order + 1);
  mxArray* wL = emx_out(len, nsh, cfg.basis); mxArray* wR = emx_out(len, nsh, cfg.basis);
  emx_check(emagls_design_emagls(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                 mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetScalar(prhs[4]),
                                 mxGetDoubles(prhs[5]), mxGetDoubles(prhs[6]), M, order, mxGetScalar(prhs[8]), len,
                                 1, 1, NULL, emx_ptr(wL), emx_ptr(wR), NULL));
  emx_return2(nlhs, plhs, wL, wR);
}
