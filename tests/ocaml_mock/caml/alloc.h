#ifndef MOCK_CAML_ALLOC_H
#define MOCK_CAML_ALLOC_H
#include "mlvalues.h"
value caml_alloc_string(size_t len);
value caml_copy_int64(int64_t v);
#endif
