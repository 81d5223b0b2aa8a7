"""Host-side terrain helpers for the terrain-aware neighbours of the MDP step.

* ``terrain_column_range``      - ``_get_terrain_column_range`` (V/mdp/utils.py:16-41), evaluated once on the host
* ``is_env_assigned_to_terrain``- V/mdp/utils.py:44-70, a lookup on ``terrain_types`` evaluated once per task
* ``TerrainGridBuffers``        - the device tensors ``rl_is_robot_on_terrain`` / ``rl_command_pit_restrict`` read
* ``HeightFieldBuffers``        - the device tensors ``rl_height_scan_cast`` reads (height field + ray pattern)
* ``grid_origins`` / ``grid_pattern_ray_starts`` - the regular layouts IsaacLab's TerrainGenerator / GridPatternCfg
  produce [IL]; used by the synthetic state provider and the tests.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _native as nat
from .cfg import RayCasterCfg, TerrainCfg


def terrain_column_range(cfg: TerrainCfg, terrain_name: str) -> tuple[int, int] | None:
    """Columns [start, end) of the named sub-terrain: fp32 cumsum of the normalised proportions times ``num_cols``,
    python ``round`` (V/mdp/utils.py:16-41). ``None`` when the terrain has no such sub-terrain."""
    if cfg.terrain_type != "generator" or terrain_name not in cfg.sub_terrains:
        return None
    names = list(cfg.sub_terrains)
    p = torch.tensor([float(x) for x in cfg.proportions], dtype=torch.float32)
    p = p / p.sum()
    c = torch.cumsum(p, dim=0)
    i = names.index(terrain_name)
    col_start = round((0.0 if i == 0 else c[i - 1].item()) * cfg.num_cols)
    col_end = round(c[i].item() * cfg.num_cols)
    return col_start, col_end


def is_env_assigned_to_terrain(cfg: TerrainCfg, terrain_types: torch.Tensor, terrain_name: str) -> torch.Tensor:
    """V/mdp/utils.py:44-70: ``terrain_types`` stores the column of every env's cell. uint8 mask [N]."""
    rng = terrain_column_range(cfg, terrain_name)
    if rng is None:
        return torch.zeros_like(terrain_types, dtype=torch.uint8)
    return ((terrain_types >= rng[0]) & (terrain_types < rng[1])).to(torch.uint8)


def grid_origins(cfg: TerrainCfg) -> torch.Tensor:
    """[num_rows, num_cols, 3] cell centres of a generated terrain, centred on the world origin (TerrainGenerator
    [IL]: ((row + 0.5) * size_x - num_rows * size_x / 2, (col + 0.5) * size_y - num_cols * size_y / 2, 0))."""
    r = (torch.arange(cfg.num_rows, dtype=torch.float64) + 0.5) * cfg.size[0] - cfg.num_rows * cfg.size[0] * 0.5
    c = (torch.arange(cfg.num_cols, dtype=torch.float64) + 0.5) * cfg.size[1] - cfg.num_cols * cfg.size[1] * 0.5
    out = torch.zeros(cfg.num_rows, cfg.num_cols, 3, dtype=torch.float64)
    out[:, :, 0] = r[:, None]
    out[:, :, 1] = c[None, :]
    return out.float()


def grid_pattern_ray_starts(cfg: RayCasterCfg) -> torch.Tensor:
    """[R, 3] ray starts of the grid height scanner in the sensor frame, x fastest ("xy" ordering), with the sensor
    offset added (patterns.grid_pattern + RayCaster._initialize_rays_impl [IL]; V/velocity_env_cfg.py:70-77)."""
    x = torch.arange(start=-cfg.size[0] / 2, end=cfg.size[0] / 2 + 1.0e-9, step=cfg.resolution)
    y = torch.arange(start=-cfg.size[1] / 2, end=cfg.size[1] / 2 + 1.0e-9, step=cfg.resolution)
    gx, gy = torch.meshgrid(x, y, indexing="xy")
    starts = torch.zeros(gx.numel(), 3)
    starts[:, 0] = gx.flatten()
    starts[:, 1] = gy.flatten()
    starts[:, 2] += cfg.offset_z
    return starts


@dataclass
class TerrainGridBuffers:
    """Device copy of ``terrain.terrain_origins`` + the column range of one named sub-terrain."""

    origins: torch.Tensor          # [rows, cols, 3] fp32, contiguous, on the device
    col_start: int
    col_end: int

    @staticmethod
    def create(cfg: TerrainCfg, terrain_name: str, device, origins: torch.Tensor | None = None) -> "TerrainGridBuffers":
        o = (grid_origins(cfg) if origins is None else origins).to(device=device, dtype=torch.float32).contiguous()
        rng = terrain_column_range(cfg, terrain_name)
        cs, ce = rng if rng is not None else (0, 0)
        return TerrainGridBuffers(o, cs, ce)

    def to_ctypes(self) -> nat.RlTerrainGrid:
        rows, cols, _ = self.origins.shape
        return nat.RlTerrainGrid(self.origins.data_ptr(), rows, cols, self.col_start, self.col_end)


@dataclass
class HeightFieldBuffers:
    """Device height field (vertex heights, x-major) + ray pattern of the height scanner."""

    heights: torch.Tensor          # [num_x, num_y] fp32, contiguous, on the device
    x0: float
    y0: float
    horizontal_scale: float
    ray_starts: torch.Tensor       # [R, 3] fp32 on the device

    def to_ctypes(self) -> nat.RlHeightField:
        nx, ny = self.heights.shape
        return nat.RlHeightField(self.heights.data_ptr(), nx, ny, self.x0, self.y0, self.horizontal_scale,
                                 self.ray_starts.shape[0], self.ray_starts.data_ptr())
