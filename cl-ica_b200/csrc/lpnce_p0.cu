#define CLICA_P 0
#include "lpnce_inst.cuh"
