/* MEX drop-in for lib/getEMagLsFiltersFromAtf.m:1, binding emagls_design_from_atf().
 * [wMlsL, wMlsR] = getEMagLsFiltersFromAtf(hL, hR, hrirGridAziZenRad, atfIrs, atfGridAziZenRad, fs, filterLen, fTrans)
 * Build: mex -R2018a -I../include getEMagLsFiltersFromAtf.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 8) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLsFiltersFromAtf needs 8 arguments");
  emagls_config cfg; emagls_config_default(&cfg);
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]);
  const mwSize* ad = mxGetDimensions(prhs[3]);            /* atfIrs: [samples x mics x directions] */
  if (mxGetNumberOfDimensions(prhs[3]) != 3) mexErrMsgIdAndTxt("eMagLS:atf", "atfIrs must be samples x mics x directions");
  const int Ta = (int)ad[0], M = (int)ad[1], Da = (int)ad[2];
  const int len = (int)mxGetScalar(prhs[6]);
  double dev = 0.0;
  mxArray* wL = mxCreateDoubleMatrix(len, M, mxREAL); mxArray* wR = mxCreateDoubleMatrix(len, M, mxREAL);
  emx_check(emagls_design_from_atf(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                   mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), Ta, M, Da, mxGetDoubles(prhs[4]),
                                   mxGetScalar(prhs[5]), len, mxGetScalar(prhs[7]), mxGetDoubles(wL), mxGetDoubles(wR),
                                   NULL, &dev));
  /* the reference's disp() line, lib/getEMagLsFiltersFromAtf.m:96 */
  mexPrintf("Matching HRTF and ATF grids, average grid deviation: %g deg\n", dev);
  emx_return2(nlhs, plhs, wL, wR);
}
