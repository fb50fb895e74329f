// gdn_types.hpp -- scalar types and constants of the solver API.
// Mirrors the reference's include/common.h:35-82 (non-LONG_TYPES build) so that
// code written against BFSSolver/PRSolver/SpmvSolver compiles unchanged.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

typedef float ScoreT;      // include/common.h:42
typedef float ValueT;      // :43
typedef int DistT;         // :45
typedef int IndexT;        // :47
typedef int WeightT;       // :48
typedef int32_t VertexId;  // :60
typedef std::vector<VertexId> VertexList;

#ifndef MYINFINITY
#define MYINFINITY 1000000000  // include/common.h:66
#endif

// PageRank constants, src/pr/pr.h:5-13
#ifndef EPSILON
#define EPSILON 0.0001
#endif
static const float kDamp = 0.85f;
#ifndef MAX_ITER
#define MAX_ITER 100
#endif
