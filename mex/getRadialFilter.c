/* MEX drop-in for dependencies/getRadialFilter.m:1, binding emagls_radial_filter().
 * radFilts = getRadialFilter(params)
 * Build: mex -R2018a -I../include getRadialFilter.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs;
  if (nrhs < 1 || !mxIsStruct(prhs[0])) mexErrMsgIdAndTxt("eMagLS:nargin", "getRadialFilter needs a params struct");
  const mxArray* p = prhs[0];
  emagls_config cfg; emagls_config_default(&cfg);
  emagls_radial_params rp; emx_radial_params(p, &rp);
  const int order = (int)emx_fld(p, "order", 4);
  const int nfft = (int)(emx_fld(p, "oversamplingFactor", 2) * emx_fld(p, "irLen", 256));
  plhs[0] = mxCreateDoubleMatrix((mwSize)(nfft / 2 + 1), (mwSize)(order + 1), mxCOMPLEX);
  emx_check(emagls_radial_filter(emx_handle(), &cfg, &rp, order, emx_fld(p, "fs", 48000), emx_fld(p, "smaRadius", 0.042),
                                 nfft, (double*)mxGetComplexDoubles(plhs[0])));
}
