/* MEX drop-in for dependencies/getSMAIRMatrix.m:1, binding emagls_smair_matrix().
 * [smairMat, params] = getSMAIRMatrix(params)      (radialFilter ~= 'none' goes through emagls_smair_matrix_radial)
 * Fields read (defaults of getSMAIRMatrix.m:30-84): smaDesignAziZenRad, order, fs, smaRadius, arrayType,
 * oversamplingFactor, irLen, returnRawMicSigs, shDefinition, radialFilter.
 * Build: mex -R2018a -I../include getSMAIRMatrix.c -L../emagls_b200/lib -lemagls_cuda   (needs MATLAB) */
#include "emagls_mex_common.h"

#define fld emx_fld

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1 || !mxIsStruct(prhs[0])) mexErrMsgIdAndTxt("eMagLS:nargin", "getSMAIRMatrix needs a params struct");
  const mxArray* p = prhs[0];
  const mxArray* mics = mxGetField(p, 0, "smaDesignAziZenRad");
  if (!mics) mexErrMsgIdAndTxt("eMagLS:fallback", "default microphone layout: use the reference getSMAIRMatrix.m");
  emagls_config cfg; emagls_config_default(&cfg);
  char buf[32] = "rigid";
  const mxArray* f = mxGetField(p, 0, "arrayType");
  if (f) mxGetString(f, buf, sizeof buf);
  cfg.array_type = strcmp(buf, "open") == 0 ? EMAGLS_ARRAY_OPEN : EMAGLS_ARRAY_RIGID;
  f = mxGetField(p, 0, "shDefinition");
  if (f) { mxGetString(f, buf, sizeof buf); cfg.basis = strcmp(buf, "complex") == 0 ? EMAGLS_BASIS_COMPLEX : EMAGLS_BASIS_REAL; }
  const int raw = (int)fld(p, "returnRawMicSigs", 0);
  f = mxGetField(p, 0, "radialFilter");
  int radial = 0;
  emagls_radial_params rp; emagls_radial_params_default(&rp);
  if (!raw) {
    /* default 'regul' is not a getRadialFilter kind: the reference errors there too (getRadialFilter.m:65) */
    if (!f) mexErrMsgIdAndTxt("eMagLS:radialFilter", "Unkown radialFilter parameter \"regul\".");
    emx_radial_params(p, &rp);
    radial = rp.kind != EMAGLS_RADIAL_NONE;
  }
  const int M = (int)mxGetM(mics), order = (int)fld(p, "order", 4);
  const double fs = fld(p, "fs", 48000), r = fld(p, "smaRadius", 0.042);
  const int nfft = (int)(fld(p, "oversamplingFactor", 4) * fld(p, "irLen", 2048));
  int simN = 0;
  emx_check(emagls_smair_matrix(emx_handle(), &cfg, mxGetDoubles(mics), mxGetDoubles(mics) + M, M, order, fs, r, nfft,
                                raw, NULL, &simN));
  const mwSize dims[3] = {(mwSize)(raw ? M : (order + 1) * (order + 1)), (mwSize)((simN + 1) * (simN + 1)),
                          (mwSize)(nfft / 2 + 1)};
  plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxCOMPLEX);
  if (radial)
    emx_check(emagls_smair_matrix_radial(g_handle, &cfg, &rp, mxGetDoubles(mics), mxGetDoubles(mics) + M, M, order, fs,
                                         r, nfft, (double*)mxGetComplexDoubles(plhs[0]), &simN));
  else
    emx_check(emagls_smair_matrix(g_handle, &cfg, mxGetDoubles(mics), mxGetDoubles(mics) + M, M, order, fs, r, nfft,
                                  raw, (double*)mxGetComplexDoubles(plhs[0]), &simN));
  if (nlhs > 1) plhs[1] = mxDuplicateArray(p);   /* the reference echoes params with defaults filled */
}
