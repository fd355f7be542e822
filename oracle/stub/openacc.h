/* Host shim so the reference's openacc.cpp compiles with g++ (its OpenACC pragmas are ignored).
 * Only the two runtime calls it makes (openacc.cpp:64-68) are declared; both are no-ops on the host. */
#ifndef FDTD_ORACLE_OPENACC_SHIM_H
#define FDTD_ORACLE_OPENACC_SHIM_H
typedef enum { acc_device_nvidia = 4 } acc_device_t;
static inline void acc_init(acc_device_t) {}
static inline void acc_set_device_num(int, acc_device_t) {}
#endif
