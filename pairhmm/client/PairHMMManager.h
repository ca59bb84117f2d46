// PairHMMManager.h -- starts the in-process accelerator manager for the PairHMM path when the host program has not
// configured one itself.  The reference's bench builds its manager from a conf file (host/main.cpp:252-273,
// pairhmm/xlnx.conf); pairhmm_manager_from_conf() does the same with this repo's pairhmm/cuda.conf, and
// pairhmm_default_manager() serves "PairHMM" on every visible GPU with the task plugin found next to this library
// (or named by $PAIRHMM_TASK_LIB).  Both publish the manager on port 1027, where PairHMMClient looks.
#ifndef PAIRHMM_MANAGER_H
#define PAIRHMM_MANAGER_H
#include <string>

#include "blaze/PlatformManager.h"

blaze::PlatformManager* pairhmm_default_manager(const char* plugin_path = nullptr, int slots_per_device = 4);
blaze::PlatformManager* pairhmm_manager_from_conf(const std::string& conf_path);
void pairhmm_shutdown_manager();
std::string pairhmm_default_plugin_path();

#endif
