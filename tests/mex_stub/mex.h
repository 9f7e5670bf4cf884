/* Minimal declarations of the MATLAB MEX API used by the gateways in mex/, for a syntax/signature check only
 * (`gcc -fsyntax-only`, tests/test_cabi.py).  Not the MathWorks header; nothing is linked against it. */
#ifndef MEX_STUB_H
#define MEX_STUB_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef struct { double real, imag; } mxComplexDouble;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
typedef int bool_t;
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
double mxGetScalar(const mxArray*);
double* mxGetDoubles(const mxArray*);
mxComplexDouble* mxGetComplexDoubles(const mxArray*);
int mxIsEmpty(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsStruct(const mxArray*);
int mxIsLogicalScalarTrue(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxDuplicateArray(const mxArray*);
mxArray* mxCreateString(const char*);
void mxDestroyArray(mxArray*);
int mexCallMATLAB(int, mxArray*[], int, mxArray*[], const char*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexPrintf(const char*, ...);
int mexAtExit(void (*)(void));
#endif
