"""In-memory trajectory input (reference mdsuite/file_io/script_input.py:8-45)."""
from typing import Iterator

from mdsuite_b200.database.simulation_database import TrajectoryChunkData, TrajectoryMetadata
from mdsuite_b200.file_io.file_read import FileProcessor


class ScriptInput(FileProcessor):
    def __init__(self, data: TrajectoryChunkData, metadata: TrajectoryMetadata, name: str):
        self.data = data
        self.mdata = metadata
        self.name = name
        if self.data.chunk_size != self.mdata.n_configurations:
            raise ValueError("data must hold all configurations described by the metadata")

    def __str__(self):
        return self.name

    def _get_metadata(self) -> TrajectoryMetadata:
        return self.mdata

    def get_configurations_generator(self) -> Iterator[TrajectoryChunkData]:
        yield self.data
