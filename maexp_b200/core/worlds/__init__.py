"""`generate_world` (core/worlds/__init__.py:19-28).  Agent state lives on the GPU; the world
object only carries the map and the agent count (core/worlds/default_world.py:17-20)."""
from ..utils.logging import initialize_logging

logger = initialize_logging(__name__)


class AbstractWorld:
    def __init__(self, config, world_map):
        self.config = config
        self.map = world_map
        self.num_agents = int(config.num_agents)


class DefaultWorld(AbstractWorld):
    pass


def generate_world(config, world_map):
    if config.world == "default":
        return DefaultWorld(config=config, world_map=world_map)
    logger.warning(f"Unexpected world is given. config.world: {config.world}")
    raise ValueError()
