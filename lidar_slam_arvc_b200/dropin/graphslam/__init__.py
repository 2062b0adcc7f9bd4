"""Drop-in `graphslam` package: only `loopclosing` is replaced.  Every other sub-module
(`graphslam.graphSLAM`, which needs gtsam) keeps resolving to the reference tree further down sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
