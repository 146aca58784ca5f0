// TEST INFRASTRUCTURE (oracle/): plain (untraced) build of the reference driver.
#include <vector>
#include "Utils.hpp"
void coalref_trace_open(const char*, long, const std::vector<AncestralTypeMaps>*) {}
void coalref_trace_close() {}
