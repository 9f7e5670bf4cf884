/* MEX drop-in for dependencies/applyRadialFilter.m:1, binding emagls_apply_radial_filter().
 * sigFiltered = applyRadialFilter(inSig, params)
 * Build: mex -R2018a -I../include applyRadialFilter.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs;
  if (nrhs < 2 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("eMagLS:nargin", "applyRadialFilter(inSig, params)");
  const mxArray* p = prhs[1];
  emagls_config cfg; emagls_config_default(&cfg);
  emagls_radial_params rp; emx_radial_params(p, &rp);
  const int order = (int)emx_fld(p, "order", 4), nfft = (int)emx_fld(p, "nfft", 0);
  const long long n = (long long)mxGetM(prhs[0]);
  if ((int)mxGetN(prhs[0]) != (order + 1) * (order + 1))
    mexErrMsgIdAndTxt("eMagLS:size", "inSig must have (order+1)^2 columns");
  if (n < nfft) mexPrintf("applyRadialFilter: short signal, applying zero padding!\n");   /* applyRadialFilter.m:25 */
  plhs[0] = mxCreateDoubleMatrix((mwSize)emagls_apply_radial_filter_rows(n, nfft), mxGetN(prhs[0]), mxREAL);
  emx_check(emagls_apply_radial_filter(emx_handle(), &cfg, &rp, mxGetDoubles(prhs[0]), n, order, emx_fld(p, "fs", 48000),
                                       emx_fld(p, "smaRadius", 0.042), nfft, mxGetDoubles(plhs[0])));
}
