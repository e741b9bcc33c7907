#define CLICA_P 1
#include "lpnce_inst.cuh"
