#ifndef MOCK_CAML_FAIL_H
#define MOCK_CAML_FAIL_H
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
void caml_failwith(const char *msg) __attribute__((noreturn));
#endif
