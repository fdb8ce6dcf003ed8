"""File-reader interface (reference mdsuite/file_io/file_read.py:35-78)."""
import abc
from typing import Iterator

from mdsuite_b200.database.simulation_database import TrajectoryChunkData, TrajectoryMetadata


class FileProcessor(abc.ABC):
    """Anything that can describe a trajectory and stream it in chunks of configurations."""

    @abc.abstractmethod
    def __str__(self) -> str:
        """A unique name (the file path); used for the ``read_files`` de-duplication."""

    @property
    def metadata(self) -> TrajectoryMetadata:
        if getattr(self, "_mdata", None) is None:
            self._mdata = self._get_metadata()
        return self._mdata

    @abc.abstractmethod
    def _get_metadata(self) -> TrajectoryMetadata:
        ...

    @abc.abstractmethod
    def get_configurations_generator(self) -> Iterator[TrajectoryChunkData]:
        ...
