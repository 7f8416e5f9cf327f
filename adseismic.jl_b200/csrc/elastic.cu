// elastic.cu -- placeholder; replaced by the real implementation.
#include "common.cuh"
#define NOTYET(name) { adseis_set_error(name ": elastic path not built yet"); return ADSEIS_ESTATE; }
ADSEIS_API int adseis_elastic_plan_create(adseis_ctx*, const adseis_elastic_params*, const adseis_slab*, int64_t, const int64_t*, const int64_t*, const int64_t*, int64_t, const int64_t*, const int64_t*, const int64_t*, size_t, adseis_elastic_plan**) NOTYET("elastic_plan_create")
ADSEIS_API int adseis_elastic_plan_destroy(adseis_elastic_plan*) { return ADSEIS_OK; }
ADSEIS_API int adseis_elastic_plan_set_model(adseis_elastic_plan*, const double*, const double*, const double*, int) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_set_srcv(adseis_elastic_plan*, const double*, int64_t, int) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_set_obs(adseis_elastic_plan*, const double*, int) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_forward(adseis_elastic_plan*) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_gradient(adseis_elastic_plan*, int) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_get(adseis_elastic_plan*, int, double*, int) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_get_snapshot(adseis_elastic_plan*, int, int64_t, double*, int) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_info(adseis_elastic_plan*, int64_t*) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_ipc_export(adseis_elastic_plan*, void*) NOTYET("elastic")
ADSEIS_API int adseis_elastic_plan_ipc_connect(adseis_elastic_plan*, const void*, const void*) NOTYET("elastic")
ADSEIS_API int adseis_elastic_forward(adseis_ctx*, const adseis_elastic_params*, const double*, const double*, const double*, int64_t, const int64_t*, const int64_t*, const int64_t*, const double*, int64_t, int64_t, const int64_t*, const int64_t*, const int64_t*, double*, double*) NOTYET("elastic")
ADSEIS_API int adseis_elastic_misfit_grad(adseis_ctx*, const adseis_elastic_params*, const double*, const double*, const double*, int64_t, const int64_t*, const int64_t*, const int64_t*, const double*, int64_t, int64_t, const int64_t*, const int64_t*, const int64_t*, const double*, double*, double*, double*, double*, double*, double*) NOTYET("elastic")
