/* MEX drop-in for lib/getEMagLs2Filters.m (same name, same positional arguments, placed earlier on
 * the MATLAB path).  Binds emagls_design_emagls2() of libemagls_cuda (include/emagls_cuda.h).
 * Build where MATLAB exists:  mex -R2018a -I../include getEMagLs2Filters.c -L../emagls_b200/lib -lemagls_cuda
 * NOTE: cannot be compiled or exercised in this repo's build image (no MATLAB, no mex.h); the C ABI it
 * binds is verified through the Python ctypes binding instead.
 *
 * [wMlsL, wMlsR] = getEMagLs2Filters(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, ...
 *                       micGridAziRad, micGridZenRad, order, fs, len, shDefinition, shFunction)
 */
#include <string.h>
#include "mex.h"
#include "emagls_cuda.h"

static emagls_handle g_handle = NULL;
static void at_exit(void) { if (g_handle) { emagls_destroy(g_handle); g_handle = NULL; } }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 10) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLs2Filters needs at least 10 arguments");
  if (nrhs >= 12 && !mxIsEmpty(prhs[11])) {
    /* a non-default shFunction cannot be evaluated on the device */
    mxArray* name; mexCallMATLAB(1, &name, 1, (mxArray**)&prhs[11], "func2str");
    char buf[64]; mxGetString(name, buf, sizeof buf);
    if (strcmp(buf, "getSH") != 0) mexErrMsgIdAndTxt("eMagLS:shFunction", "only @getSH is supported by the CUDA drop-in");
  }
  emagls_config cfg; emagls_config_default(&cfg);
  if (nrhs >= 11 && !mxIsEmpty(prhs[10])) {
    char def[16]; mxGetString(prhs[10], def, sizeof def);
    cfg.basis = strcmp(def, "complex") == 0 ? EMAGLS_BASIS_COMPLEX : EMAGLS_BASIS_REAL;
  }
  if (!g_handle) {
    if (emagls_create(0, &g_handle) != EMAGLS_OK) mexErrMsgIdAndTxt("eMagLS:cuda", "no usable CUDA device");
    mexAtExit(at_exit);
  }
  const int T = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]);
  const int M = (int)mxGetNumberOfElements(prhs[5]);
  const int order = (int)mxGetScalar(prhs[7]), len = (int)mxGetScalar(prhs[9]);
  plhs[0] = mxCreateDoubleMatrix(len, M, mxREAL);
  mxArray* wR = mxCreateDoubleMatrix(len, M, mxREAL);
  int rc = emagls_design_emagls2(g_handle, &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                 mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetScalar(prhs[4]),
                                 mxGetDoubles(prhs[5]), mxGetDoubles(prhs[6]), M, order, mxGetScalar(prhs[8]), len,
                                 1, 1, NULL, mxGetDoubles(plhs[0]), mxGetDoubles(wR), NULL);
  if (rc != EMAGLS_OK) mexErrMsgIdAndTxt("eMagLS:cuda", "%s", emagls_last_error(g_handle)); /* e.g. 'len too short' */
  if (nlhs > 1) plhs[1] = wR; else mxDestroyArray(wR);
}
