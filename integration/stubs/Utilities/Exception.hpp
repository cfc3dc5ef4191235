// STUB of src/Utilities/Exception.hpp:10-16 (Throw).
#pragma once
#include <stdexcept>
template <class E> [[noreturn]] void Throw(const E& e) { throw e; }
