#ifndef MDS_CONFIG_H
#define MDS_CONFIG_H
#define MDS_SET_MAX 256
#define MDS_ID_TYPE int
#endif
