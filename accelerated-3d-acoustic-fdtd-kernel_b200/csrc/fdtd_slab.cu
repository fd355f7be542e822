// fdtd_slab.cu -- x-slab neighbours (multi-GPU halo exchange).  Filled in by the slab milestone.
#include "fdtd_plan.h"

extern "C" int fdtd_b200_plan_ipc_export(fdtd_b200_plan *, void *) { return (int)cudaErrorNotSupported; }
extern "C" int fdtd_b200_plan_ipc_attach(fdtd_b200_plan *, int, const void *) { return (int)cudaErrorNotSupported; }
extern "C" int fdtd_b200_plan_attach_local(fdtd_b200_plan *, int, fdtd_b200_plan *) { return (int)cudaErrorNotSupported; }
