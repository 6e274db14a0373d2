"""TEST INFRASTRUCTURE: the folder table the GGUF nodes register their file lists in."""
folder_names_and_paths = {"diffusion_models": (["/models/diffusion_models"], {".safetensors"}), "text_encoders": (["/models/text_encoders"], {".safetensors"})}


def get_filename_list(key):
    return []


def get_full_path(key, name):
    return f"/models/{key}/{name}"


def get_folder_paths(key):
    return [f"/models/{key}"]
