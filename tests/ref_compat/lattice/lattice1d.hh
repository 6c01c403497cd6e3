// the reference header name, served by the drop-in host layer (host/mgmc_host.hh)
#include "mgmc_host.hh"
