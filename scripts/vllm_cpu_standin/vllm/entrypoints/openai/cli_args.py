def make_arg_parser(parser):
    parser.add_argument("--model", type=str, default=None)
    parser.add_argument("--port", type=int, default=8000)
    parser.add_argument("--host", type=str, default=None)
    parser.add_argument("--enable-sleep-mode", action="store_true")
    parser.add_argument("--tensor-parallel-size", type=int, default=1)
    return parser


def validate_parsed_serve_args(args):
    return None
