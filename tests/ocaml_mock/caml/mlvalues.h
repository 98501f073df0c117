/* Minimal stand-in for the OCaml runtime headers: just enough declarations for
 * `gcc -fsyntax-only ocaml/zkb200_stubs.c` (tests/test_cpu_contract.py).  NOT the OCaml runtime. */
#ifndef MOCK_CAML_MLVALUES_H
#define MOCK_CAML_MLVALUES_H
#include <stddef.h>
#include <stdint.h>
typedef intptr_t value;
#define CAMLprim
#define Val_unit ((value)1)
#define Int_val(v) ((int)((v) >> 1))
#define Long_val(v) ((long)((v) >> 1))
#define Field(v, i) (((value *)(v))[i])
#define Wosize_val(v) ((size_t)(((value *)(v))[-1] >> 10))
#define Bytes_val(v) ((unsigned char *)(v))
#define Int64_val(v) (*(int64_t *)(((value *)(v)) + 1))
size_t caml_string_length(value v);
#endif
