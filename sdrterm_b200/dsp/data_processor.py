"""Processor interface (reference: src/dsp/data_processor.py:23-41)."""
from abc import ABC, abstractmethod


class DataProcessor(ABC):
    """A consumer of IQ chunks."""

    @abstractmethod
    def processData(self, *args, **kwargs) -> None:
        """Consume chunks from a queue until ``isDead`` is set or the end-of-stream marker arrives,
        writing the demodulated float64 frames to ``f`` (a file name; None = stdout).

        isDead: multiprocessing.Value-like flag; buffer: queue of chunks; f: output file name.
        """
        pass
