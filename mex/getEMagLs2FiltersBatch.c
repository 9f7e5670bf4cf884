/* Batched extension of the getEMagLs2Filters drop-in (NOT in the reference API): one call designs the filter
 * banks of a whole head-orientation grid and / or several HRTF sets, which is where the CUDA path's throughput
 * comes from (a loop of 3600 single calls pays the per-call setup 3600 times).
 *
 * [wMlsL, wMlsR] = getEMagLs2FiltersBatch(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, ...
 *                       micGridAziRad, micGridZenRad, order, fs, len, rotations, applyDiffusenessConst)
 *   hL, hR     [numSamples x numDirections x numSets]
 *   rotations  [3 x 3 x B]: head orientation b sees HRIR-grid direction u at R(:,:,b) * u, i.e. page b equals one
 *              reference call with hrirGridAziRad/ZenRad rotated by R(:,:,b)  (lib/getEMagLs2Filters.m:1-2)
 *   applyDiffusenessConst  optional, default false: the diffuse-field covariance constraint of earlier reference
 *              versions (CHANGELOG.md:10-18), an extension of this library (emagls_config.diffuseness_const)
 *   wMlsL/R    [len x numMics x (numSets * B)], set index slowest
 * Binds emagls_design_emagls2() with num_sets / num_orient / rotations (include/emagls_cuda.h).
 * Build where MATLAB exists:  mex -R2018a -I../include getEMagLs2FiltersBatch.c -L../emagls_b200/lib -lemagls_cuda
 * NOTE: cannot be compiled or exercised in this repo's build image (no MATLAB, no mex.h).
 */
#include <stdlib.h>
#include "emagls_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 11) mexErrMsgIdAndTxt("eMagLS:nargin", "getEMagLs2FiltersBatch needs 11 arguments");
  emagls_config cfg; emagls_config_default(&cfg);
  if (nrhs > 11 && !mxIsEmpty(prhs[11]) && mxGetScalar(prhs[11]) != 0.0) cfg.diffuseness_const = 1;
  const mwSize* hd = mxGetDimensions(prhs[0]);
  const int T = (int)hd[0], D = (int)hd[1];
  const int sets = mxGetNumberOfDimensions(prhs[0]) > 2 ? (int)hd[2] : 1;
  const int M = (int)mxGetNumberOfElements(prhs[5]);
  const int order = (int)mxGetScalar(prhs[7]), len = (int)mxGetScalar(prhs[9]);
  const mwSize* rd = mxGetDimensions(prhs[10]);
  if (rd[0] != 3 || rd[1] != 3) mexErrMsgIdAndTxt("eMagLS:rotations", "rotations must be [3 x 3 x B]");
  const int B = mxGetNumberOfDimensions(prhs[10]) > 2 ? (int)rd[2] : 1;
  /* MATLAB pages are column-major 3x3; the C ABI takes row-major 3x3 per orientation */
  const double* rm = mxGetDoubles(prhs[10]);
  double* rot = (double*)malloc((size_t)B * 9 * sizeof(double));
  int b, i, j;
  for (b = 0; b < B; ++b)
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) rot[b * 9 + i * 3 + j] = rm[b * 9 + j * 3 + i];
  mwSize od[3]; od[0] = (mwSize)len; od[1] = (mwSize)M; od[2] = (mwSize)sets * (mwSize)B;
  mxArray* wL = mxCreateNumericArray(3, od, mxDOUBLE_CLASS, mxREAL);
  mxArray* wR = mxCreateNumericArray(3, od, mxDOUBLE_CLASS, mxREAL);
  const int rc = emagls_design_emagls2(emx_handle(), &cfg, mxGetDoubles(prhs[0]), mxGetDoubles(prhs[1]), T, D,
                                       mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetScalar(prhs[4]),
                                       mxGetDoubles(prhs[5]), mxGetDoubles(prhs[6]), M, order, mxGetScalar(prhs[8]), len,
                                       sets, B, rot, mxGetDoubles(wL), mxGetDoubles(wR), NULL);
  free(rot);
  emx_check(rc);
  emx_return2(nlhs, plhs, wL, wR);
}
