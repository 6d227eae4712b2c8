// oracle/shim — TEST INFRASTRUCTURE: FluidObject with the accessor surface the solver uses
// (GetPositions/GetPositionCount/GetVelocity — VFD/Source/Simulation/DFSPH/FluidObject/FluidObject.h:22-32)
// but constructed from raw positions, so the oracle and the CUDA path get the *same* initial
// state without going through the mesh sampler (scene preparation is outside the hot path).
#ifndef VFD_ORACLE_SHIM_FLUID_OBJECT_H
#define VFD_ORACLE_SHIM_FLUID_OBJECT_H
#include "pch.h"
namespace vfd {
    struct FluidObject : public RefCounted {
        FluidObject(const std::vector<glm::vec3>& positions, const glm::vec3& velocity) : m_Positions(positions), m_Velocity(velocity) {}
        const std::vector<glm::vec3>& GetPositions() const { return m_Positions; }
        unsigned int GetPositionCount() const { return (unsigned int)m_Positions.size(); }
        const glm::vec3& GetVelocity() const { return m_Velocity; }
    private:
        std::vector<glm::vec3> m_Positions;
        glm::vec3 m_Velocity;
    };
}
#endif
