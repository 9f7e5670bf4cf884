/* MEX drop-in for dependencies/binauralDecode.m binding emagls_binaural_decode().
 * binauralOut = binauralDecode(in, inFs, decL, decR, decFs, compensateDelay)
 * (resampling, mono-signal convolution and horRotAngleRad fall through to the original .m file)
 * Build: mex -R2018a -I../include binauralDecode.c -L../emagls_b200/lib -lemagls_cuda  (needs MATLAB). */
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 5) mexErrMsgIdAndTxt("eMagLS:nargin", "binauralDecode needs at least 5 arguments");
  if (nrhs > 6 || mxGetScalar(prhs[1]) != mxGetScalar(prhs[4]) || mxIsComplex(prhs[0]) || mxIsComplex(prhs[2])) {
    mexErrMsgIdAndTxt("eMagLS:fallback", "use the reference binauralDecode.m for this argument combination");
  }
  const long long n = (long long)mxGetM(prhs[0]);
  const int ch = (int)mxGetN(prhs[0]), len = (int)mxGetM(prhs[2]);
  const int comp = (nrhs > 5) && mxIsLogicalScalarTrue(prhs[5]);
  const long long rows = comp ? n - len / 2 + 1 : n;
  plhs[0] = mxCreateDoubleMatrix((mwSize)rows, 2, mxREAL);
  emx_check(emagls_binaural_decode(emx_handle(), mxGetDoubles(prhs[0]), n, ch, mxGetDoubles(prhs[2]),
                                  mxGetDoubles(prhs[3]), len, comp, mxGetDoubles(plhs[0])));
}
