/*
 * cgo_result.h — the {result, error-string} pair every C entry point returns.
 * Byte-compatible with the reference's cgoutils/utils.h:20-23; Go side:
 * cgoutils/utils.go:25-33 (DoCGoCall: if pStrErr != nil -> GoString, C.free, panic).
 */
#ifndef ARESDB_B200_CGO_RESULT_H_
#define ARESDB_B200_CGO_RESULT_H_

typedef struct {
  void *res;           /* integer result cast to pointer, or an allocated pointer        */
  const char *pStrErr; /* NULL on success; malloc'd message owned (freed) by the caller  */
} CGoCallResHandle;

#endif /* ARESDB_B200_CGO_RESULT_H_ */
