/* MEX drop-in for lib/getMagLsArrayDiffuseFilter.m:1, binding emagls_array_diffuse_filter().
 * wAdf = getMagLsArrayDiffuseFilter(micRadius, micGridAziRad, micGridZenRad, order, fs, len, shDefinition, shFunction)
 * Build: mex -R2018a -I../include getMagLsArrayDiffuseFilter.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs;
  if (nrhs < 6) mexErrMsgIdAndTxt("eMagLS:nargin", "getMagLsArrayDiffuseFilter needs at least 6 arguments");
  emx_require_default_handle(nrhs, prhs, 7, "getSH");
  emagls_config cfg; emagls_config_default(&cfg);
  cfg.basis = emx_basis(nrhs, prhs, 6);
  const int M = (int)mxGetNumberOfElements(prhs[1]), len = (int)mxGetScalar(prhs[5]);
  plhs[0] = mxCreateDoubleMatrix((mwSize)len, 1, mxREAL);
  emx_check(emagls_array_diffuse_filter(emx_handle(), &cfg, mxGetScalar(prhs[0]), mxGetDoubles(prhs[1]),
                                        mxGetDoubles(prhs[2]), M, (int)mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), len,
                                        mxGetDoubles(plhs[0])));
}
