/* schema_l7.h — placeholder until l7.cu lands (same commit series). */
#pragma once
#define L7_COLUMNS(X) X(_ID, "_id", DFGPU_CH_UINT64, 8)
