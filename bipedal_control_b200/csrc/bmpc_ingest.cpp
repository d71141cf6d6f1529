// Readers for the reference's own configuration files (Boost property-tree INFO subset + URDF subset) and the
// model reduction of centroidal_model::createPinocchioInterface [UPSTREAM] (BipedalRobotInterface.cpp:117).
#include <stdexcept>
#include "bmpc_model.h"
namespace bmpc {
HostModel load_reference_files(const std::string&, const std::string&, const std::string&, const std::string&) {
  throw std::invalid_argument("[bmpc] INFO/URDF ingestion not built yet: pass bmpc_config.model_file");
}
}  // namespace bmpc
