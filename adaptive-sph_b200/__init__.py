"""adaptive-sph_b200 — B200-native step loop for kaegi/adaptive-sph (hot path only).

The directory name carries a hyphen (it is the name the task fixes); import it with
`importlib.import_module("adaptive-sph_b200")` or through `asph_b200.py` at the repo root.
"""
from .binding import (AsphError, FluidSimulation, StatisticsRecorder, init_fluid_sim, load_library, FIELDS,
                      LEVEL_INTERIOR, PRODUCT_LIB)
from .params import SimulationParams, merge_overwrite
from .scene import (SceneConfig, add_fluid_block, init_simulation_params, scene_boundary, scene_particle_count,
                    scene_particles)
from .distributed import (DistributedFluidSimulation, broadcast_unique_id, gather_by_global_index, owner_of, share_range,
                          slab_bounds_from_histogram)
from .split_patterns import SplitPatterns, load_split_patterns_from_file
from .export_jobs import ImageExportConfig, JobError, export_simulation_image, load_job_file, resolve_job, run_job
from .vtk import VtkExporter, init_fluid_sim_from_vtk, read_vtk_file, write_vtk_file, write_vtk_file2

__all__ = ["AsphError", "FluidSimulation", "StatisticsRecorder", "init_fluid_sim", "load_library", "FIELDS",
           "LEVEL_INTERIOR", "PRODUCT_LIB", "SimulationParams", "merge_overwrite", "SceneConfig", "add_fluid_block",
           "init_simulation_params", "scene_boundary", "scene_particles", "scene_particle_count", "SplitPatterns",
           "load_split_patterns_from_file", "DistributedFluidSimulation", "broadcast_unique_id", "gather_by_global_index",
           "owner_of", "share_range", "slab_bounds_from_histogram", "VtkExporter", "init_fluid_sim_from_vtk", "read_vtk_file",
           "write_vtk_file", "write_vtk_file2", "ImageExportConfig", "JobError", "export_simulation_image", "load_job_file",
           "resolve_job", "run_job"]
