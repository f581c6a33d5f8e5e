from .base import ProjectionEntry, ProjectionOperator, create_projection_map, project, register
from . import clamp, box, cone, simplex  # noqa: F401  (registers "box", "cone", "simplex", "simplex_eq")

__all__ = ["project", "register", "ProjectionOperator", "create_projection_map", "ProjectionEntry"]
