// Internal (non-exported) helpers shared between the translation units of libomc.
#pragma once
int omc_sm_count();

// Prior / generic precision-matrix kinds understood by the small-matrix kernels.
//   0: scaled identity  (ptr may be null => 1.0, else ptr[0] is the diagonal value)
//   1: diagonal vector  (p entries)
//   2: dense row-major  (p x p)
enum OmcMatKind { OMC_MAT_EYE = 0, OMC_MAT_DIAG = 1, OMC_MAT_DENSE = 2 };
