/* MEX drop-in for lib/getMagLsSphericalHeadFilter.m:1, binding emagls_spherical_head_filter().
 * [wShf, W_Shf] = getMagLsSphericalHeadFilter(micRadius, order, fs, len)
 * Build: mex -R2018a -I../include getMagLsSphericalHeadFilter.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 4) mexErrMsgIdAndTxt("eMagLS:nargin", "getMagLsSphericalHeadFilter(micRadius, order, fs, len)");
  emagls_config cfg; emagls_config_default(&cfg);
  const int len = (int)mxGetScalar(prhs[3]);
  const int nfft = 2 * len < cfg.nfft_max_len ? 2 * len : cfg.nfft_max_len;
  mxArray* w = mxCreateDoubleMatrix((mwSize)len, 1, mxREAL);
  mxArray* W = mxCreateDoubleMatrix((mwSize)nfft, 1, mxREAL);
  emx_check(emagls_spherical_head_filter(emx_handle(), &cfg, mxGetScalar(prhs[0]), (int)mxGetScalar(prhs[1]),
                                         mxGetScalar(prhs[2]), len, mxGetDoubles(w), mxGetDoubles(W)));
  emx_return2(nlhs, plhs, w, W);
}
