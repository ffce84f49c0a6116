#include "Kokkos_Version_Info.hpp"
namespace Kokkos { namespace Impl {
std::string GIT_BRANCH = "oracle";
std::string GIT_COMMIT_HASH = "94dd7e0e";
std::string GIT_CLEAN_STATUS = "CLEAN";
std::string GIT_COMMIT_DESCRIPTION = "";
std::string GIT_COMMIT_DATE = "";
} }
