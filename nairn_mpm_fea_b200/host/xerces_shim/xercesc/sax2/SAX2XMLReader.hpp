// oracle-only Xerces stand-in: everything lives in util/PlatformUtils.hpp
#include <xercesc/util/PlatformUtils.hpp>
