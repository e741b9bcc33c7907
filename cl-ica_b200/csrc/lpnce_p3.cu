#define CLICA_P 3
#include "lpnce_inst.cuh"
