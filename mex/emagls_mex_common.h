/* Shared helpers of the MEX drop-ins (interleaved-complex API: build with `mex -R2018a`).
 * NOTE: none of the gateways can be compiled in this repo's build image (no MATLAB, no mex.h); the C ABI
 * they bind is exercised through the Python ctypes binding. */
#ifndef EMAGLS_MEX_COMMON_H
#define EMAGLS_MEX_COMMON_H
#include <string.h>
#include "mex.h"
#include "emagls_cuda.h"

static emagls_handle g_handle = NULL;
static void emx_at_exit(void) { if (g_handle) { emagls_destroy(g_handle); g_handle = NULL; } }

static emagls_handle emx_handle(void) {
  if (!g_handle) {
    if (emagls_create(0, &g_handle) != EMAGLS_OK) mexErrMsgIdAndTxt("eMagLS:cuda", "no usable CUDA device");
    mexAtExit(emx_at_exit);
  }
  return g_handle;
}

/* shDefinition argument ('real' | 'complex', default 'real') */
static int emx_basis(int nrhs, const mxArray* prhs[], int idx) {
  char def[16] = "real";
  if (nrhs > idx && !mxIsEmpty(prhs[idx])) mxGetString(prhs[idx], def, sizeof def);
  return strcmp(def, "complex") == 0 ? EMAGLS_BASIS_COMPLEX : EMAGLS_BASIS_REAL;
}

/* shFunction / chFunction handles: only the defaults can be evaluated on the device */
static void emx_require_default_handle(int nrhs, const mxArray* prhs[], int idx, const char* expected) {
  if (nrhs > idx && !mxIsEmpty(prhs[idx])) {
    mxArray* name; char buf[64];
    mexCallMATLAB(1, &name, 1, (mxArray**)&prhs[idx], "func2str");
    mxGetString(name, buf, sizeof buf);
    if (strcmp(buf, expected) != 0)
      mexErrMsgIdAndTxt("eMagLS:handle", "only @%s is supported by the CUDA drop-in", expected);
  }
}

/* 1 if prhs[idx] is a function handle other than @<expected> (then the shim evaluates it on the host) */
static int emx_is_custom_handle(int nrhs, const mxArray* prhs[], int idx, const char* expected) {
  if (nrhs > idx && !mxIsEmpty(prhs[idx])) {
    mxArray* name; char buf[64];
    mexCallMATLAB(1, &name, 1, (mxArray**)&prhs[idx], "func2str");
    mxGetString(name, buf, sizeof buf);
    mxDestroyArray(name);
    return strcmp(buf, expected) != 0;
  }
  return 0;
}

/* Y = shFunction(order, [azi zen], 'real') through feval: [numel(azi) x (order+1)^2] real */
static mxArray* emx_eval_basis(const mxArray* fn, int order, const double* azi, const double* zen, mwSize n) {
  mxArray* args[4]; mxArray* out = NULL;
  mwSize i;
  args[0] = (mxArray*)fn;
  args[1] = mxCreateDoubleMatrix(1, 1, mxREAL); mxGetDoubles(args[1])[0] = (double)order;
  args[2] = mxCreateDoubleMatrix(n, 2, mxREAL);
  for (i = 0; i < n; ++i) { mxGetDoubles(args[2])[i] = azi[i]; mxGetDoubles(args[2])[n + i] = zen[i]; }
  args[3] = mxCreateString("real");
  if (mexCallMATLAB(1, &out, 4, args, "feval") != 0 || mxIsComplex(out))
    mexErrMsgIdAndTxt("eMagLS:handle", "shFunction must return a real [directions x (order+1)^2] matrix");
  mxDestroyArray(args[1]); mxDestroyArray(args[2]); mxDestroyArray(args[3]);
  return out;
}

static mxArray* emx_out(mwSize rows, mwSize cols, int basis) {
  return mxCreateDoubleMatrix(rows, cols, basis == EMAGLS_BASIS_COMPLEX ? mxCOMPLEX : mxREAL);
}
static double* emx_ptr(mxArray* a) {
  return mxIsComplex(a) ? (double*)mxGetComplexDoubles(a) : mxGetDoubles(a);
}
static void emx_check(int rc) {
  if (rc != EMAGLS_OK) mexErrMsgIdAndTxt("eMagLS:cuda", "%s", emagls_last_error(g_handle)); /* e.g. 'len too short' */
}
static void emx_return2(int nlhs, mxArray* plhs[], mxArray* wL, mxArray* wR) {
  plhs[0] = wL;
  if (nlhs > 1) plhs[1] = wR; else mxDestroyArray(wR);
}

/* struct field helpers + the getRadialFilter parameter block (getRadialFilter.m:9-23, 48-56) */
static double emx_fld(const mxArray* s, const char* name, double dflt) {
  const mxArray* f = mxGetField(s, 0, name);
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}
static void emx_radial_params(const mxArray* p, emagls_radial_params* rp) {
  char buf[32] = "tikhonov";
  const mxArray* f = mxGetField(p, 0, "radialFilter");
  emagls_radial_params_default(rp);
  if (f) mxGetString(f, buf, sizeof buf);
  if (strcmp(buf, "none") == 0) rp->kind = EMAGLS_RADIAL_NONE;
  else if (strcmp(buf, "tikhonov") == 0) rp->kind = EMAGLS_RADIAL_TIKHONOV;
  else if (strcmp(buf, "softlimit") == 0) rp->kind = EMAGLS_RADIAL_SOFTLIMIT;
  else if (strcmp(buf, "full") == 0) rp->kind = EMAGLS_RADIAL_FULL;
  else mexErrMsgIdAndTxt("eMagLS:radialFilter", "Unkown radialFilter parameter \"%s\".", buf);
  rp->regul_const = emx_fld(p, "regulConst", 1e-2);
  rp->noise_gain_db = emx_fld(p, "noiseGainDb", 20.0);
  strcpy(buf, "rigid");
  f = mxGetField(p, 0, "arrayType");
  if (f) mxGetString(f, buf, sizeof buf);
  rp->array_type = strcmp(buf, "open") == 0 ? EMAGLS_ARRAY_OPEN : EMAGLS_ARRAY_RIGID;
  strcpy(buf, "planeWave");
  f = mxGetField(p, 0, "waveModel");
  if (f) mxGetString(f, buf, sizeof buf);
  if (rp->kind != EMAGLS_RADIAL_NONE && strcmp(buf, "pointSource") == 0)
    mexErrMsgIdAndTxt("eMagLS:waveModel", "WaveModel parameter \"%s\" not yet implemented.", buf);
}
#endif
