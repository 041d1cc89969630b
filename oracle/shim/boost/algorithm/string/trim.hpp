// TEST INFRASTRUCTURE ONLY (oracle/)
#include "../string.hpp"
