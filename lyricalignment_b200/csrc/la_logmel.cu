// la_logmel.cu -- K1 placeholder until the tcgen05 front end lands.
#include "../../include/lyricalign.h"
extern "C" {
size_t la_logmel_workspace_bytes(int, int64_t) { return 0; }
int la_logmel(const float*, int, int64_t, int64_t, float*, int64_t, void*, void*) { return LA_ERR_ARG; }
}
