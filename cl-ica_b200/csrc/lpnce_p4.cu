#define CLICA_P 4
#include "lpnce_inst.cuh"
