/* MEX drop-in for lib/getEMagLs2Filters.m (same name, same positional arguments, placed earlier on
 * the MATLAB path).  Binds emagls_design_emagls2() of libemagls_cuda (include/emagls_cuda.h); a shFunction
 * handle other than @getSH is evaluated here on the host and the two basis matrices go down through
 * emagls_design_sma_basis() (example handle at verifyEMagLs.m:356-368).
 * Build where MATLAB exists:  mex -R2018a -I../include getEMagLs2Filters.c -L../emagls_b200/lib -lemagls_cuda
 * NOTE: cannot be compiled or exercised in this repo's build image (no MATLAB, no mex.h); the C ABI it
 * binds is verified through the Python ctypes binding instead.
 *
 * [wMlsL, wMlsR] = getEMagLs2Filters(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, ...
 *                       micGridAziRad, micGridZenRad, order, fs, len, shDefinition, shFunction)
 */
#include <math.h>
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 10) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLs2Filters needs at least 10 arguments");
  emagls_config cfg; emagls_config_default(&cfg);
  cfg.basis = emx_basis(nrhs, prhs, 10);                 /* eMagLS2 outputs are real for either basis */
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]);
  const int M = (int)mxGetNumberOfElements(prhs[5]);
  const int order = (int)mxGetScalar(prhs[7]), len = (int)mxGetScalar(prhs[9]);
  const double r = mxGetScalar(prhs[4]), fs = mxGetScalar(prhs[8]);
  mxArray* wL = mxCreateDoubleMatrix(len, M, mxREAL); mxArray* wR = mxCreateDoubleMatrix(len, M, mxREAL);
  if (emx_is_custom_handle(nrhs, prhs, 11, "getSH")) {
    /* simulation order of getSMAIRMatrix.m:95 */
    int simN = (int)ceil(fs * 3.141592653589793 * r / cfg.speed_of_sound);
    if (simN < order) simN = order;
    if (cfg.basis != EMAGLS_BASIS_REAL)
      mexErrMsgIdAndTxt("eMagLS:handle", "custom shFunction handles are supported for shDefinition = 'real'");
    mxArray* Yh = emx_eval_basis(prhs[11], simN, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), (mwSize)D);
    mxArray* Ym = emx_eval_basis(prhs[11], simN, mxGetDoubles(prhs[5]), mxGetDoubles(prhs[6]), (mwSize)M);
    emx_check(emagls_design_sma_basis(emx_handle(), &cfg, 0, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                      mxGetDoubles(Yh), (simN + 1) * (simN + 1), r, mxGetDoubles(Ym), M, order, fs, len,
                                      1, 1, mxGetDoubles(wL), mxGetDoubles(wR), NULL));
    mxDestroyArray(Yh); mxDestroyArray(Ym);
  } else {
    emx_check(emagls_design_emagls2(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                    mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), r, mxGetDoubles(prhs[5]),
                                    mxGetDoubles(prhs[6]), M, order, fs, len, 1, 1, NULL, mxGetDoubles(wL),
                                    mxGetDoubles(wR), NULL));
  }
  emx_return2(nlhs, plhs, wL, wR);
}
