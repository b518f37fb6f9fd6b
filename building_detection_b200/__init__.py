"""building-detection B200: sm_100a implementation of the reference's ensemble inference hot path
(tiled forward of five segmentation networks -> 3-of-5 mask fusion -> contour extraction) behind the
reference's own Python call surface.  See DESIGN.md."""
__version__ = "0.1.0"
