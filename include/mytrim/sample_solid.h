// sample_solid.h — kept so that sources written against the reference's header layout compile unchanged;
// the whole plugin surface lives in mytrim.h.
#ifndef MYTRIM_B200_FWD_SAMPLE_SOLID_H
#define MYTRIM_B200_FWD_SAMPLE_SOLID_H
#include "mytrim.h"
#endif
