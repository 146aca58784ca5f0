// Test infrastructure (oracle/): the reference includes this Boost header only for the transitive
// standard headers it drags in (Utils.hpp:43,90; TypeContainer.hpp:36,124); its own VectorHash
// (Utils.hpp:60-76) does the hashing.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <map>
#include <algorithm>
#include <iostream>
#include <string>
