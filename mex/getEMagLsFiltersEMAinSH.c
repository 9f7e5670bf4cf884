/* MEX drop-in for lib/getEMagLsFiltersEMAinSH.m:1-2, binding emagls_design_ema_sh().
 * [wMlsL, wMlsR] = getEMagLsFiltersEMAinSH(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, micGridAziRad,
 *                                          order, fs, len, shDefinition, shFunction, chFunction)
 * Build: mex -R2018a -I../include getEMagLsFiltersEMAinSH.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 9) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLsFiltersEMAinSH needs at least 9 arguments");
  emx_require_default_handle(nrhs, prhs, 10, "getSH");
  emx_require_default_handle(nrhs, prhs, 11, "getCH");
  emagls_config cfg; emagls_config_default(&cfg);
  cfg.basis = emx_basis(nrhs, prhs, 9);
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]), M = (int)mxGetNumberOfElements(prhs[5]);
  const int order = (int)mxGetScalar(prhs[6]), len = (int)mxGetScalar(prhs[8]), nch = (order + 1) * (order + 1);
  mxArray* wL = emx_out(len, nch, cfg.basis); mxArray* wR = emx_out(len, nch, cfg.basis);
  emx_check(emagls_design_ema_sh(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D, mxGetDoubles(prhs[2]),
            mxGetDoubles(prhs[3]), mxGetScalar(prhs[4]), mxGetDoubles(prhs[5]), M, order, mxGetScalar(prhs[7]), len,
            emx_ptr(wL), emx_ptr(wR), NULL));
  emx_return2(nlhs, plhs, wL, wR);
}
