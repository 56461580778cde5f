"""Caffe interoperability: protobuf wire codec for the messages Znicz consumes."""
from .protobuf2 import Datum, BlobProto, BlobShape, decode_message, parse_net_text  # noqa
