#ifndef SCOREC_CONFIG_H
#define SCOREC_CONFIG_H
#define SCOREC_NO_MPI
#endif
