#ifndef MOCK_CAML_MEMORY_H
#define MOCK_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam1(a) (void)(a)
#define CAMLparam3(a, b, c) (void)(a), (void)(b), (void)(c)
#define CAMLparam2(a, b) (void)(a); (void)(b)
#define CAMLparam4(a, b, c, d) (void)(a); (void)(b); (void)(c); (void)(d)
#define CAMLparam5(a, b, c, d, e) (void)(a); (void)(b); (void)(c); (void)(d); (void)(e)
#define CAMLxparam1(a) (void)(a)
#define CAMLlocal1(a) value a = 0
#define CAMLreturn(x) return (x)
#endif
