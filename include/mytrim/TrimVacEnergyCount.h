// TrimVacEnergyCount.h — kept so that sources written against the reference's header layout compile unchanged;
// the whole plugin surface lives in mytrim.h.
#ifndef MYTRIM_B200_FWD_TRIMVACENERGYCOUNT_H
#define MYTRIM_B200_FWD_TRIMVACENERGYCOUNT_H
#include "mytrim.h"
#endif
