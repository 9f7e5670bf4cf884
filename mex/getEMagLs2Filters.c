/* MEX drop-in for lib/getEMagLs2Filters.m (same name, same positional arguments, placed earlier on
 * the MATLAB path).  Binds emagls_design_emagls2() of libemagls_cuda (include/emagls_cuda.h).
 * Build where MATLAB exists:  mex -R2018a -I../include getEMagLs2Filters.c -L../emagls_b200/lib -lemagls_cuda
 * NOTE: cannot be compiled or exercised in this repo's build image (no MATLAB, no mex.h); the C ABI it
 * binds is verified through the Python ctypes binding instead.
 *
 * [wMlsL, wMlsR] = getEMagLs2Filters(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, ...
 *                       micGridAziRad, micGridZenRad, order, fs, len, shDefinition, shFunction)
 */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 10) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLs2Filters needs at least 10 arguments");
  emx_require_default_handle(nrhs, prhs, 11, "getSH");   /* a custom shFunction cannot run on the device */
  emagls_config cfg; emagls_config_default(&cfg);
  cfg.basis = emx_basis(nrhs, prhs, 10);                 /* eMagLS2 outputs are real for either basis */
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]);
  const int M = (int)mxGetNumberOfElements(prhs[5]);
  const int order = (int)mxGetScalar(prhs[7]), len = (int)mxGetScalar(prhs[9]);
  mxArray* wL = mxCreateDoubleMatrix(len, M, mxREAL); mxArray* wR = mxCreateDoubleMatrix(len, M, mxREAL);
  emx_check(emagls_design_emagls2(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                  mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetScalar(prhs[4]),
                                  mxGetDoubles(prhs[5]), mxGetDoubles(prhs[6]), M, order, mxGetScalar(prhs[8]), len,
                                  1, 1, NULL, mxGetDoubles(wL), mxGetDoubles(wR), NULL));
  emx_return2(nlhs, plhs, wL, wR);
}
