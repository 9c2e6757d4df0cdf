# Drop-in fall-through: a module this package does not provide (the reference's out-of-scope parts -- title generation,
# training datasets, language-model utilities) resolves to the reference's own file when the reference's
# video_chapter_generation/ directory sits BEHIND this package on sys.path (route A).
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
