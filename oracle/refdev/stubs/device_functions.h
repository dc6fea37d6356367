// oracle/refdev/stubs/device_functions.h -- empty stand-in.  The reference's include/util/dmath.h includes CUDA's
// internal <device_functions.h>, which cannot be compiled by a host compiler; nothing of it is needed on the host.
