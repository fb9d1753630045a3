from .shared_buffer import SharedReplayBuffer  # noqa: F401
