"""Stand-in for NVIDIA Isaac Gym Preview 4 (closed source, not installable here).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference package
under /root/reference can be imported in the build container and driven with
synthetic simulator tensors.  Nothing in the shipped
package ``ti5_isaacgym_b200`` imports it.
"""
from . import gymapi, gymtorch, gymutil, torch_utils, terrain_utils  # noqa: F401
