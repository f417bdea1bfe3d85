from typing import TypeVar

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
