from .path_utils import ensure_directory_exists, find_files, is_valid_file  # noqa: F401
