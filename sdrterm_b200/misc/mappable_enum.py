"""``MappableEnum``: an Enum whose members can be listed as a name -> value mapping (reference:
src/misc/mappable_enum.py:23-30; the CLI builds its ``--encoding`` choices from it)."""
from __future__ import annotations

from enum import Enum


class MappableEnum(Enum):

    @classmethod
    def dict(cls) -> dict:
        return {member.name: member.value for member in cls}

    @classmethod
    def tuples(cls) -> tuple:
        return tuple((member.name, member.value) for member in cls)
