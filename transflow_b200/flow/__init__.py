from .sources.source import FlowSource

Direction = FlowSource.Direction
LockMode = FlowSource.LockMode
