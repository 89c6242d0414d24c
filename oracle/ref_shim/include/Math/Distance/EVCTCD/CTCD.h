// shadows Library/Math/Distance/EVCTCD/CTCD.h (exact CTCD back-end; never called on the hot path)
#pragma once
