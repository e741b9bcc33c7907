#define CLICA_P 5
#include "lpnce_inst.cuh"
