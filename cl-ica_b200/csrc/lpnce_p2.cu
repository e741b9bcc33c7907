#define CLICA_P 2
#include "lpnce_inst.cuh"
