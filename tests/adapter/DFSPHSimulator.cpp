// tests/adapter/DFSPHSimulator.cpp — the adapter of INTEGRATION.md section 2, compiled.
//
// A replacement BODY for the reference's VFD/Source/Simulation/DFSPH/DFSPHSimulator.cpp: every member function that
// DFSPHSimulator.h (:10-56) declares, implemented over the C ABI of libvfd_dfsph.so (include/vfd_dfsph.h) instead of
// DFSPHImplementation.  The reference's header is used AS IT IS (this file compiles against it where it lies under
// /root/reference), so nothing above DFSPHSimulation — ComponentPanel, TimelinePanel, Scene::OnRender — changes; the
// state the header has no member for (the library handle, the copies behind the reference-returning getters) lives in
// a side table keyed by the object.  A maintainer who may edit the header would turn the table into members.
//
// Scene preparation stays reference code: FluidObject samples the mesh, RigidBody builds the SDF and the volume map
// (RigidBody.cu:10-73); the adapter flattens each body's map the way SDF::GetDeviceData does before its upload
// (Utility/SDF/SDF.cu:227-306).  The SDF keeps those vectors private (SDF.cuh:39-41): VFD_ADAPTER_OPEN_SDF opens them for
// this translation unit (the test build); upstream would add a three-line accessor.
//
// VFD_ADAPTER_HEADLESS (the test build): no Application / ThreadPool / SystemInfo — Simulate() runs on the calling thread.
#include "pch.h"
#ifdef VFD_ADAPTER_OPEN_SDF
#define private public
#include "Utility/SDF/SDF.cuh"
#include "Simulation/DFSPH/RigidBody/RigidBody.cuh"
#undef private
#endif
#include "Simulation/DFSPH/DFSPHSimulator.h"
#ifndef VFD_ADAPTER_HEADLESS
#include "Core/Application.h"
#include "Debug/SystemInfo.h"
#endif
#include "vfd_dfsph.h"

#include <mutex>
#include <unordered_map>

namespace vfd
{
	namespace
	{
		// what DFSPHImplementation held and the header's getters hand out by reference
		struct AdapterState
		{
			VfdDfsph* Handle = nullptr;
			DFSPHSimulationDescription Description;
			DFSPHSimulationInfo Info;
			DFSPHDebugInfo DebugInfo;
			PrecomputedDFSPHCubicKernel Kernel;
			ParticleSearch Search;                               // inspector only (GetByteSize / GetBounds): never searched
			std::vector<Ref<RigidBody>> RigidBodies;
			Ref<DFSPHParticleBuffer> FrameBuffer;
			Ref<VertexArray> NoVertexArray;
		};

		std::mutex s_TableMutex;
		std::unordered_map<const DFSPHSimulation*, AdapterState*> s_Table;

		AdapterState& StateOf(const DFSPHSimulation* self)
		{
			std::lock_guard<std::mutex> lock(s_TableMutex);
			return *s_Table.at(self);
		}

		VfdDfsphDescription ToAbi(const DFSPHSimulationDescription& d)
		{
			VfdDfsphDescription o;
			vfd_dfsph_default_description(&o);
			o.TimeStepSize = d.TimeStepSize; o.MinTimeStepSize = d.MinTimeStepSize; o.MaxTimeStepSize = d.MaxTimeStepSize;
			o.FrameLength = d.FrameLength; o.FrameCount = d.FrameCount;
			o.MinPressureSolverIterations = d.MinPressureSolverIterations; o.MaxPressureSolverIterations = d.MaxPressureSolverIterations;
			o.MaxPressureSolverError = d.MaxPressureSolverError;
			o.EnableDivergenceSolverError = d.EnableDivergenceSolverError;
			o.MinDivergenceSolverIterations = d.MinDivergenceSolverIterations; o.MaxDivergenceSolverIterations = d.MaxDivergenceSolverIterations;
			o.MaxDivergenceSolverError = d.MaxDivergenceSolverError;
			o.EnableViscositySolver = d.EnableViscositySolver;
			o.MinViscositySolverIterations = d.MinViscositySolverIterations; o.MaxViscositySolverIterations = d.MaxViscositySolverIterations;
			o.MaxViscositySolverError = d.MaxViscositySolverError;
			o.Viscosity = d.Viscosity; o.BoundaryViscosity = d.BoundaryViscosity; o.TangentialDistanceFactor = d.TangentialDistanceFactor;
			o.EnableSurfaceTensionSolver = d.EnableSurfaceTensionSolver; o.SurfaceTensionSmoothPassCount = d.SurfaceTensionSmoothPassCount;
			o.SurfaceTension = d.SurfaceTension; o.TemporalSmoothing = d.TemporalSmoothing;
			o.CSDFix = d.CSDFix; o.CSD = d.CSD;
			o.ParticleRadius = d.ParticleRadius;
			o.Gravity[0] = d.Gravity.x; o.Gravity[1] = d.Gravity.y; o.Gravity[2] = d.Gravity.z;
			return o;
		}

		// the library reports instead of exiting (the reference: print + exit(1), Compute/Utility/CUDA/cutil.h:772-778)
		void Check(VfdDfsph* handle, int status, const char* what)
		{
			if (status != VFD_OK) {
				fprintf(stderr, "[vfd_dfsph] %s failed (%d): %s\n", what, status, handle ? vfd_dfsph_last_error(handle) : "no handle");
			}
		}

		// DFSPHSimulationInfo is the 128-byte struct the ABI mirrors field for field (include/vfd_dfsph.h: VfdDfsphInfo)
		void RefreshInfo(AdapterState& s)
		{
			static_assert(sizeof(VfdDfsphInfo) == sizeof(DFSPHSimulationInfo), "DFSPHSimulationInfo layout");
			VfdDfsphInfo info;
			vfd_dfsph_get_info(s.Handle, &info);
			memcpy(&s.Info, &info, sizeof info);
		}

		void RefreshDebugInfo(AdapterState& s)
		{
			VfdDfsphDebugInfo d;
			vfd_dfsph_get_debug_info(s.Handle, &d);
			s.DebugInfo.IterationCount = d.IterationCount;
			s.DebugInfo.DivergenceSolverIterationCount = d.DivergenceSolverIterationCount; s.DebugInfo.DivergenceSolverError = d.DivergenceSolverError;
			s.DebugInfo.PressureSolverIterationCount = d.PressureSolverIterationCount; s.DebugInfo.PressureSolverError = d.PressureSolverError;
			s.DebugInfo.ViscositySolverIterationCount = d.ViscositySolverIterationCount; s.DebugInfo.ViscositySolverError = d.ViscositySolverError;
			s.DebugInfo.FrameTime = d.FrameTime; s.DebugInfo.FrameIndex = d.FrameIndex;
		}
	}

	DFSPHSimulation::DFSPHSimulation(const DFSPHSimulationDescription& desc)
	{
#ifndef VFD_ADAPTER_HEADLESS
		if (SystemInfo::CUDADeviceMeetsRequirements() == false) {
			return;
		}
#endif
		AdapterState* state = new AdapterState();
		state->Description = desc;
		const VfdDfsphDescription abi = ToAbi(desc);
		const int status = vfd_dfsph_create(&abi, 0, &state->Handle);      // the reference drives device 0 (Debug/SystemInfo.cpp:34-35)
		Check(state->Handle, status, "vfd_dfsph_create");
		{
			std::lock_guard<std::mutex> lock(s_TableMutex);
			s_Table[this] = state;
		}
		if (status != VFD_OK) {
			return;
		}
		state->Kernel.SetRadius(4.0f * desc.ParticleRadius);              // DFSPHImplementation::SetDescription (:328-331)
		RefreshInfo(*state);
		m_Initialized = true;
	}

	DFSPHSimulation::~DFSPHSimulation()
	{
		AdapterState* state = nullptr;
		{
			std::lock_guard<std::mutex> lock(s_TableMutex);
			auto it = s_Table.find(this);
			if (it != s_Table.end()) { state = it->second; s_Table.erase(it); }
		}
		if (state) {
			if (state->Handle) vfd_dfsph_destroy(state->Handle);
			delete state;
		}
	}

	const Ref<VertexArray>& DFSPHSimulation::GetVertexArray()
	{
		AdapterState& s = StateOf(this);
		return s.FrameBuffer ? s.FrameBuffer->GetVertexArray() : s.NoVertexArray;      // DFSPHImplementation::GetVertexArray (:283)
	}

	// FluidObject::GetPositions() is the sampler's output, one velocity per object (FluidObject.cpp:27-40,
	// DFSPHImplementation::SetFluidObjects :172-253)
	void DFSPHSimulation::SetFluidObjects(const std::vector<Ref<FluidObject>>& fluidObjects)
	{
		AdapterState& s = StateOf(this);
		std::vector<float> positions, velocities;
		for (const Ref<FluidObject>& fluid : fluidObjects)
		{
			const glm::vec3 v = fluid->GetVelocity();
			for (const glm::vec3& p : fluid->GetPositions())
			{
				positions.insert(positions.end(), { p.x, p.y, p.z });
				velocities.insert(velocities.end(), { v.x, v.y, v.z });
			}
		}
		const uint32_t count = static_cast<uint32_t>(positions.size() / 3u);
		Check(s.Handle, vfd_dfsph_set_particles(s.Handle, positions.data(), velocities.data(), count), "vfd_dfsph_set_particles");
		RefreshInfo(s);
		// created here, on the caller's (UI) thread: the buffer owns GL objects (DFSPHImplementation.cu:250)
		s.FrameBuffer = Ref<DFSPHParticleBuffer>::Create(s.Description.FrameCount, count);
	}

	// One VfdVolumeMap per body: the arrays SDF::GetDeviceData flattens (SDF.cu:234-283), still on the host
	void DFSPHSimulation::SetRigidBodies(const std::vector<Ref<RigidBody>>& rigidBodies)
	{
		AdapterState& s = StateOf(this);
		s.RigidBodies = rigidBodies;
		const size_t bodyCount = rigidBodies.size();
		std::vector<VfdVolumeMap> maps(bodyCount);
		std::vector<std::vector<float>> nodes(bodyCount);
		std::vector<std::vector<uint32_t>> cells(bodyCount), cellMaps(bodyCount);
		for (size_t b = 0; b < bodyCount; b++)
		{
#ifdef VFD_ADAPTER_OPEN_SDF
			const SDF& sdf = *rigidBodies[b]->m_DensityMap.Raw();
			VfdVolumeMap& m = maps[b];
			for (int k = 0; k < 3; k++)
			{
				m.domainMin[k] = sdf.m_Domain.min[k]; m.domainMax[k] = sdf.m_Domain.max[k];
				m.resolution[k] = sdf.m_Resolution[k];
				m.cellSize[k] = sdf.m_CellSize[k]; m.cellSizeInverse[k] = sdf.m_CellSizeInverse[k];
			}
			m.fieldCount = static_cast<uint32_t>(sdf.m_FieldCount);
			m.nodeCount = static_cast<uint32_t>(sdf.m_Nodes[0].size());
			m.cellCount = static_cast<uint32_t>(sdf.m_CellCount);
			m.cellMapCount = static_cast<uint32_t>(sdf.m_CellMap[0].size());
			for (const auto& field : sdf.m_Nodes) nodes[b].insert(nodes[b].end(), field.begin(), field.end());
			for (const auto& field : sdf.m_Cells) for (const auto& cell : field) cells[b].insert(cells[b].end(), cell.begin(), cell.end());
			for (const auto& field : sdf.m_CellMap) cellMaps[b].insert(cellMaps[b].end(), field.begin(), field.end());
			m.nodes = nodes[b].data(); m.cells = cells[b].data(); m.cellMap = cellMaps[b].data();
#else
#error "give SDF a host accessor for m_Nodes / m_Cells / m_CellMap (SDF.cuh:39-41), or build with VFD_ADAPTER_OPEN_SDF"
#endif
		}
		Check(s.Handle, vfd_dfsph_set_rigid_bodies(s.Handle, static_cast<uint32_t>(bodyCount), maps.data()), "vfd_dfsph_set_rigid_bodies");
		RefreshInfo(s);
	}

	DFSPHImplementation::SimulationState DFSPHSimulation::GetSimulationState() const
	{
		return static_cast<DFSPHImplementation::SimulationState>(vfd_dfsph_get_state(StateOf(this).Handle));     // None / Simulating / Ready
	}

	unsigned int DFSPHSimulation::GetParticleCount() { return vfd_dfsph_get_particle_count(StateOf(this).Handle); }
	float DFSPHSimulation::GetParticleRadius() const { return vfd_dfsph_get_particle_radius(StateOf(this).Handle); }
	float DFSPHSimulation::GetMaxVelocityMagnitude() const { return vfd_dfsph_get_max_velocity_magnitude(StateOf(this).Handle); }
	float DFSPHSimulation::GetCurrentTimeStepSize() const { return vfd_dfsph_get_current_time_step_size(StateOf(this).Handle); }
	unsigned int DFSPHSimulation::GetRigidBodyCount() const { return vfd_dfsph_get_rigid_body_count(StateOf(this).Handle); }

	const ParticleSearch& DFSPHSimulation::GetParticleSearch() const { return StateOf(this).Search; }
	const DFSPHSimulationDescription& DFSPHSimulation::GetDescription() const { return StateOf(this).Description; }
	PrecomputedDFSPHCubicKernel& DFSPHSimulation::GetKernel() { return StateOf(this).Kernel; }
	const std::vector<Ref<RigidBody>>& DFSPHSimulation::GetRigidBodies() const { return StateOf(this).RigidBodies; }
	Ref<DFSPHParticleBuffer> DFSPHSimulation::GetParticleFrameBuffer() { return StateOf(this).FrameBuffer; }

	void DFSPHSimulation::SetDescription(const DFSPHSimulationDescription& desc)
	{
		AdapterState& s = StateOf(this);
		s.Description = desc;
		const VfdDfsphDescription abi = ToAbi(desc);
		Check(s.Handle, vfd_dfsph_set_description(s.Handle, &abi), "vfd_dfsph_set_description");
		s.Kernel.SetRadius(4.0f * desc.ParticleRadius);
		RefreshInfo(s);
	}

	const DFSPHSimulationInfo& DFSPHSimulation::GetInfo() const
	{
		AdapterState& s = StateOf(this);
		RefreshInfo(s);
		return s.Info;
	}

	const DFSPHDebugInfo& DFSPHSimulation::GetDebugInfo() const
	{
		AdapterState& s = StateOf(this);
		RefreshDebugInfo(s);                                             // a snapshot: safe while the bake runs on the worker
		return s.DebugInfo;
	}

	bool& DFSPHSimulation::GetRenderParticles() { return m_RenderParticles; }
	bool& DFSPHSimulation::GetRenderFlowLines() { return m_RenderFlowLines; }
	unsigned int DFSPHSimulation::GetFlowLineSampleCount() const { return m_FlowLineSampleCount; }
	const std::vector<unsigned int>& DFSPHSimulation::GetFlowLineIndices() const { return m_FlowLineIndices; }

	// evenly spaced particle indices: frames come back in ORIGINAL particle order, so an index means the same particle in
	// every frame (Scene.cpp:361-368 draws the lines through them)
	void DFSPHSimulation::RecomputeFlowLineIndices()
	{
		m_FlowLineIndices.resize(m_FlowLineSampleCount);
		const float stride = static_cast<float>(GetParticleCount()) / static_cast<float>(m_FlowLineSampleCount);
		float at = 0.0f;
		for (unsigned int& index : m_FlowLineIndices)
		{
			index = static_cast<unsigned int>(at);
			at += stride;
		}
	}

	void DFSPHSimulation::SetFlowLineCount(unsigned int count)
	{
		if (m_FlowLineSampleCount != count)
		{
			m_FlowLineSampleCount = std::min(count, GetParticleCount());
			RecomputeFlowLineIndices();
		}
	}

	void DFSPHSimulation::Simulate()
	{
		auto bake = [this]
		{
			AdapterState& s = StateOf(this);
			Check(s.Handle, vfd_dfsph_simulate(s.Handle), "vfd_dfsph_simulate");
			// the baked frames go into the reference's own frame cache: TimelinePanel and Scene::OnRender keep reading it
			uint32_t baked = 0;
			vfd_dfsph_get_frame_count(s.Handle, &baked);
			const uint32_t count = vfd_dfsph_get_particle_count(s.Handle);
			static_assert(sizeof(DFSPHParticleSimple) == sizeof(VfdParticleSimple), "DFSPHParticleSimple is the 36-byte frame record");
			for (uint32_t i = 0; i < baked && i < s.Description.FrameCount; i++)
			{
				Ref<DFSPHParticleFrame> frame = Ref<DFSPHParticleFrame>::Create();
				frame->ParticleData.resize(count);
				vfd_dfsph_get_frame(s.Handle, i, reinterpret_cast<VfdParticleSimple*>(frame->ParticleData.data()), &frame->MaxVelocityMagnitude, &frame->CurrentTimeStep);
				s.FrameBuffer->SetFrameData(i, frame);                        // DFSPHParticleBuffer.cu:26-33
			}
			m_FlowLineSampleCount = glm::min(m_FlowLineSampleCount, count);
			RecomputeFlowLineIndices();
		};
#ifdef VFD_ADAPTER_HEADLESS
		bake();
#else
		Application::Get().GetThreadPool()->PushTask(bake);                  // one worker, as in the reference (DFSPHSimulator.cpp:156-166)
#endif
	}
}
