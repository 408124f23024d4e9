#pragma once
#include <boost/multi_index_container.hpp>
