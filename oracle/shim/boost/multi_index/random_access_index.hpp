// Test infrastructure only: stand-in header (see oracle/README.md).
#pragma once
#include <boost/multi_index_container.hpp>
